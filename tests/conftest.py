import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


# a dead peer must fail a slab test, not hang the GPU box (the library's default is to wait forever)
os.environ.setdefault("SVL_SPIN_TIMEOUT_MS", "20000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _parse_meta(text):
    """Fixture metadata is the repr of a dict of plain values; `inf` / `nan` appear bare."""
    import ast
    import re
    text = re.sub(r"(?<![\w.'\"])(-?)inf(?![\w'\"])", r"\g<1>1e999", text)     # literal_eval reads 1e999 as float inf
    text = re.sub(r"(?<![\w.'\"])nan(?![\w'\"])", "None", text)
    return ast.literal_eval(text)


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    if "meta" in d:
        d["meta"] = _parse_meta(str(d["meta"]))
    return d


def golden_inputs(d, dtype=None):
    """-> (mt, eps) from a fixture (None / scalar 1.0 when absent)."""
    mt = d["mt"] if "mt" in d else None
    if "eps" in d and ("eps_is_field" not in d or bool(d["eps_is_field"])) and d["eps"].size > 1:
        eps = d["eps"]
    elif "eps" in d:
        eps = d["eps"].reshape(-1)[0]
    else:
        eps = d.get("meta", {}).get("linear_coefficient", 1.0)
    return mt, eps


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
