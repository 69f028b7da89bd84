"""pycuda.compiler.SourceModule: nvcc -> cubin -> cuModuleLoadData; functions are called like pyCUDA's
(numpy scalars by value, arrays by pointer, grid= / block= keywords) on the null stream."""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np
from cuda.bindings import driver as drv

from . import LAUNCH_COUNTS
from .driver import check

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
INCLUDE = os.path.join(HERE, "..", "include")              # stand-in pycuda-complex.hpp
PREBUILT = os.path.join(ROOT, "baseline", "_ref", "cubin")    # cubins compiled ahead (baseline/install_ref.py)
CACHE = os.path.join(tempfile.gettempdir(), "svirl_bref_cubin")


def cubin_for(code, out_dir=None):
    """pyCUDA wraps the source in extern "C" { } and runs `nvcc --cubin -arch sm_XX`; same here for sm_100a.
    (-std=c++14 instead of the reference's -std=c++11 option: the thrust header behind the pycuda::complex
    stand-in needs it.)  Cached by content hash; a cubin prebuilt in the build container is used when present."""
    h = hashlib.sha1(code.encode()).hexdigest()[:16]
    for d in (PREBUILT, CACHE):
        p = os.path.join(d, "ref_%s.cubin" % h)
        if os.path.exists(p):
            return p
    out_dir = out_dir or CACHE
    os.makedirs(out_dir, exist_ok=True)
    p = os.path.join(out_dir, "ref_%s.cubin" % h)
    src = os.path.join(out_dir, "ref_%s.cu" % h)
    with open(src, "w") as f:
        f.write('extern "C" {\n%s\n}\n' % code)
    try:
        subprocess.check_call(["nvcc", "-std=c++14", "-arch=sm_100a", "-cubin", "-w", "-I", INCLUDE, src, "-o", p + ".tmp"])
        os.replace(p + ".tmp", p)
    finally:
        os.remove(src)
    return p


class _Function(object):
    def __init__(self, handle, name):
        self.handle, self.name = handle, name

    def __call__(self, *args, **kw):
        grid = tuple(kw.get("grid", (1, 1, 1))) + (1, 1)
        block = tuple(kw.get("block", (1, 1, 1))) + (1, 1)
        # one zero-padded 16-byte slot per kernel parameter; the driver reads each parameter's declared size
        # (so np.uint32(0) / np.uintp(0) in a pointer slot both read as NULL, as with pyCUDA's packed buffer)
        n = len(args)
        slots = (ctypes.c_ubyte * (16 * n))()
        ptrs = (ctypes.c_void_p * n)()
        base = ctypes.addressof(slots)
        for k, a in enumerate(args):
            if hasattr(a, "gpudata"):
                b = np.uint64(int(a.gpudata)).tobytes()
            elif isinstance(a, (np.generic, np.ndarray)):
                b = np.ascontiguousarray(a).tobytes()
            else:
                raise TypeError("kernel argument %d of %s: %r" % (k, self.name, type(a)))
            assert len(b) <= 16, (self.name, k, len(b))
            ctypes.memmove(base + 16 * k, b, len(b))
            ptrs[k] = base + 16 * k
        LAUNCH_COUNTS[self.name] = LAUNCH_COUNTS.get(self.name, 0) + 1
        check(drv.cuLaunchKernel(self.handle, int(grid[0]), int(grid[1]), int(grid[2]), int(block[0]), int(block[1]),
                                 int(block[2]), 0, 0, ctypes.addressof(ptrs), 0))


class SourceModule(object):
    def __init__(self, code, options=None, **kw):
        with open(cubin_for(code), "rb") as f:
            self.image = f.read()
        self.module = check(drv.cuModuleLoadData(self.image))

    def get_function(self, name):
        return _Function(check(drv.cuModuleGetFunction(self.module, name.encode())), name)
