import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import build_ref  # noqa: E402
from . import LAUNCH_COUNTS  # noqa: E402
from .gpuarray import GPUArray  # noqa: E402


def _ctype(t, real_c):
    t = t.replace("const", "").strip()
    if "*" in t:
        return ctypes.c_void_p
    return {"real_t": real_c, "float": ctypes.c_float, "double": ctypes.c_double,
            "uint32_t": ctypes.c_uint32, "int32_t": ctypes.c_int32, "int": ctypes.c_int,
            "bool": ctypes.c_bool}[t]


class _Function(object):
    def __init__(self, lib, name, info, real_c):
        self.name = name
        self.fn = getattr(lib, "simt_launch_" + name)
        self.ctypes_ = [_ctype(t, real_c) for t, _ in info["params"]]
        self.fn.argtypes = [ctypes.c_int] * 4 + self.ctypes_
        self.fn.restype = None
        self.mode = 1 if info["fiber"] else 0

    def __call__(self, *args, **kw):
        grid = kw.get("grid", (1, 1, 1))
        block = kw.get("block", (1, 1, 1))
        assert len(args) == len(self.ctypes_), (self.name, len(args), len(self.ctypes_))
        conv = []
        for a, ct in zip(args, self.ctypes_):
            if ct is ctypes.c_void_p:
                if isinstance(a, GPUArray):
                    conv.append(a._a.ctypes.data)
                else:
                    assert int(a) == 0, "non-array passed as pointer"
                    conv.append(None)
            else:
                conv.append(np.asarray(a).reshape(-1)[0].item())
        LAUNCH_COUNTS[self.name] = LAUNCH_COUNTS.get(self.name, 0) + 1
        gy = grid[1] if len(grid) > 1 else 1
        self.fn(int(grid[0]), int(gy), int(block[0]), self.mode, *conv)


class SourceModule(object):
    def __init__(self, code, options=None, **kw):
        self.kernels = build_ref.parse_kernels(code)
        self.lib = ctypes.CDLL(build_ref.build_module(code))
        self.real_c = ctypes.c_float if "typedef float real_t" in code else ctypes.c_double

    def get_function(self, name):
        return _Function(self.lib, name, self.kernels[name], self.real_c)
