#!/usr/bin/env python3
"""Dev-time extractor: the reference's `#[cfg(test)]` golden vectors -> tests/golden/*.json.

The reference (Rust, cannot be built here) pins its behaviour with literal
input/output arrays inside unit tests.  This tool walks
every `#[test] fn` under /root/reference/src, collects each array literal in
source order together with the identifier it is bound to (or the text that
precedes it, for arrays inlined in `assert_eq!`), and writes one JSON file per
test.  f32 values are stored as IEEE-754 bit patterns (rounded exactly like a
rustc f32 literal) so comparisons in tests/ are bit-exact.

Scalars (e.g. `assert_eq!(result.gg_ind, 193)`) are few and are restated by hand
in the pytest files next to the call they belong to, each with its file:line.

Runs only in the development container; outputs are committed.
Usage: python tools/extract_golden.py
"""
from __future__ import annotations

import json
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).parent))
from rustlit import (balanced, f32_bits_from_float, looks_float, parse_array,  # noqa: E402
                     strip_comments)

REF = Path("/root/reference/src")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

LITERAL_BODY = re.compile(r"^[\s\d.,eE+\-_xa-fA-FtruefalsTRUEFALS]*$")


def test_functions(src: str):
    """Yield (name, body, line) for each #[test] fn (first of cfg twins wins)."""
    seen = set()
    for m in re.finditer(r"#\[test\]\s*fn\s+(\w+)\s*\(\)\s*\{", src):
        name = m.group(1)
        i = m.end()
        depth = 1
        while depth:
            c = src[i]
            depth += (c == "{") - (c == "}")
            i += 1
        if name in seen:
            continue
        seen.add(name)
        yield name, src[m.end():i - 1], src.count("\n", 0, m.start()) + 1


def arrays_in(body: str):
    i = 0
    while True:
        i = body.find("[", i)
        if i < 0:
            return
        j = balanced(body, i)
        text = body[i:j]
        inner = text[1:-1]
        if ";" in inner or "[" in inner or not LITERAL_BODY.match(inner) or not inner.strip():
            # type annotation, repeat expr, nested/indexing or non-literal: step inside
            i += 1
            continue
        n_items = len([t for t in inner.split(",") if t.strip()])
        prefix = body[max(0, i - 120):i]
        mm = re.search(r"let\s+(?:mut\s+)?(\w+)\s*(?::\s*\[([^\]]*)\])?\s*=\s*$", prefix)
        if mm:
            label, ann = mm.group(1), mm.group(2) or ""
        else:
            label = re.sub(r"\s+", " ", prefix.strip())[-60:]
            ann = ""
        # skip indexing like x[..400] / tiny shape literals handled by hand
        if re.search(r"[\w\)\]]$", body[:i].rstrip()[-1:] or " ") and not mm and n_items < 3:
            i = j
            continue
        yield label, ann, text, n_items
        i = j


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    index = {}
    for path in sorted(REF.rglob("*.rs")):
        raw = path.read_text()
        src = strip_comments(raw)
        rel = path.relative_to(REF.parent)
        for name, body, _ in test_functions(src):
            # line number in the original (un-stripped) file
            m = re.search(r"fn\s+" + name + r"\s*\(", raw)
            line = raw.count("\n", 0, m.start()) + 1
            arrays = []
            for label, ann, text, n in arrays_in(body):
                if "true" in text or "false" in text:
                    kind, vals = "bool", [int(v) for v in parse_array(text, False)]
                elif "f32" in ann or looks_float(text):
                    kind = "f32_bits"
                    vals = [f32_bits_from_float(v) for v in parse_array(text, True)]
                else:
                    kind, vals = "int", parse_array(text, False)
                arrays.append({"label": label, "kind": kind, "n": len(vals), "values": vals})
            if not arrays:
                continue
            stem = f"{rel.parent.name}__{rel.stem}__{name}"
            doc = {"source": f"{rel}:{line}", "test": name, "arrays": arrays}
            (OUT / f"{stem}.json").write_text(json.dumps(doc, separators=(",", ":")) + "\n")
            index[stem] = [(a["label"], a["kind"], a["n"]) for a in arrays]
    for k, v in index.items():
        print(k)
        for a in v:
            print("    ", a)


if __name__ == "__main__":
    main()
