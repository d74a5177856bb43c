import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(stem: str):
    """Arrays of one reference #[test] (tools/extract_golden.py), in source order.
    f32 arrays come back as np.float32 with the exact bit patterns rustc would produce."""
    doc = json.loads((GOLDEN / f"{stem}.json").read_text())
    out = []
    for a in doc["arrays"]:
        if a["kind"] == "f32_bits":
            out.append(np.array(a["values"], dtype=np.uint32).view(np.float32))
        elif a["kind"] == "bool":
            out.append(np.array(a["values"], dtype=np.uint8))
        else:
            out.append(np.array(a["values"], dtype=np.int64))
    return out


def f32(x) -> np.float32:
    """A Rust f32 decimal literal, correctly rounded (float64 -> float32 is exact enough for <= 9 digits)."""
    return np.float32(x)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
