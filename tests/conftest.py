import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    if "meta" in d:
        d["meta"] = eval(str(d["meta"]), {"inf": np.inf, "nan": np.nan})
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
