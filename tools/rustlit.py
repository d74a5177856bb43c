"""Shared helpers for the dev-time tools that read literal data out of the
reference's Rust sources (tables and golden test vectors).

Nothing here is used at run time by the product, the tests or the bench: the
tools run once in the development container (where /root/reference exists) and
their outputs are committed.
"""
from __future__ import annotations

import re
import struct
from fractions import Fraction


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return src


def f32_bits_from_float(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def f32_from_bits(b: int) -> float:
    return struct.unpack("<f", struct.pack("<I", b & 0xFFFFFFFF))[0]


def _next_f32(x: float, up: bool) -> float:
    b = f32_bits_from_float(x)
    if x == 0.0:
        return f32_from_bits(1 if up else 0x80000001)
    if (x > 0) == up:
        b += 1
    else:
        b -= 1
    return f32_from_bits(b)


def decimal_to_f32(lit: str) -> float:
    """Correctly rounded (round-half-even) decimal literal -> f32, the way
    rustc parses an f32 literal.  Going through f64 first can double-round, so
    the f64-derived candidate is checked against its neighbours exactly."""
    lit = lit.replace("_", "")
    exact = Fraction(lit)
    cand = struct.unpack("<f", struct.pack("<f", float(exact)))[0]
    best = cand
    for c in (_next_f32(cand, True), _next_f32(cand, False)):
        d_c = abs(Fraction(c) - exact)
        d_b = abs(Fraction(best) - exact)
        if d_c < d_b or (d_c == d_b and (f32_bits_from_float(c) & 1) == 0
                         and (f32_bits_from_float(best) & 1) == 1):
            best = c
    return best


_NUM = r"[-+]?(?:0x[0-9a-fA-F_]+|\d[\d_]*\.?[\d_]*(?:[eE][-+]?\d+)?)"


def parse_scalar(tok: str, is_float: bool):
    """One array element: a literal or `a / b` (the only operator the tables use)."""
    tok = tok.strip()
    tok = re.sub(r"(_?)(f32|f64|i8|i16|i32|u8|u16|u32|usize|isize)$", "", tok)
    if tok in ("true", "false"):
        return tok == "true"
    if "/" in tok:
        a, b = tok.split("/")
        if not is_float:
            return int(parse_scalar(a, False)) // int(parse_scalar(b, False))
        fa, fb = parse_scalar(a, True), parse_scalar(b, True)
        import numpy as np
        return float(np.float32(fa) / np.float32(fb))
    if not is_float:
        return int(tok.replace("_", ""), 0)
    return decimal_to_f32(tok)


def split_top(body: str) -> list[str]:
    """Split a bracket body on top-level commas."""
    out, depth, cur = [], 0, []
    for ch in body:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        out.append(tail)
    return [o.strip() for o in out if o.strip()]


def parse_array(text: str, is_float: bool):
    """`[a, b, [c, d]]` -> nested python lists; also handles `[v; n]`."""
    text = text.strip()
    assert text.startswith("[") and text.endswith("]"), text[:40]
    body = text[1:-1]
    m = re.fullmatch(r"\s*([^;\[\]]+);\s*(\d+)\s*", body)
    if m:
        return [parse_scalar(m.group(1), is_float)] * int(m.group(2))
    items = split_top(body)
    out = []
    for it in items:
        if it.startswith("["):
            out.append(parse_array(it, is_float))
        else:
            out.append(parse_scalar(it, is_float))
    return out


def balanced(src: str, start: int) -> int:
    """Index just past the bracket that closes src[start] == '['."""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "[":
            depth += 1
        elif src[i] == "]":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def looks_float(text: str) -> bool:
    body = re.sub(r"0x[0-9a-fA-F_]+", "0", text)
    return bool(re.search(r"\d\.\d|\d\.[,\s\]]|\d[eE][-+]?\d", body))
