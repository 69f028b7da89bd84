"""Scalar and vector sums on the device (API of svirl/parallel/reduction.py:10-198).
Both are deterministic two-stage reductions inside the library (svl_sum / svl_sum_v)."""
import ctypes as C

import numpy as np

import svirl_b200.config as cfg
from svirl_b200 import _lib
from svirl_b200.storage.arrays import DeviceArray


class Reduction(object):

    def __init__(self, Par):
        self.par = Par

    def gsum(self, ga_in, ga_out=None, N=0, block_size=0, use_gpuarray_sum=False):
        """Sum the first N (default: all) entries of a flat real device array.  Returns the
        value on the host, or stores it in ga_out[0] when ga_out is given."""
        if ga_in is None:
            return None
        if N == 0:
            N = ga_in.size
        out = C.c_double()
        _lib.call("svl_sum", self.par.ctx, ga_in.handle, int(N), C.byref(out))
        val = cfg.dtype(out.value)
        if ga_out is not None:
            ga_out.set(np.array([val], dtype=cfg.dtype))
            return None
        return val

    def gsum_v(self, ga_in, nv, ne, block_size=0):
        """Sum nv vectors of ne components stored back to back; returns ne host values."""
        if ga_in is None:
            return None
        if nv == 1:
            return ga_in.copy()
        out = (C.c_double * int(ne))()
        _lib.call("svl_sum_v", self.par.ctx, ga_in.handle, int(nv), int(ne), out)
        return np.array(out[:], dtype=cfg.dtype)

    # ---- test hooks used by the reference's tests/at_reduction.py
    def test_sum_v(self, a_in, nv, ne, block_size=256):
        assert ne == 5
        assert 0 <= block_size <= 1024
        ga = DeviceArray.from_host(self.par, np.ascontiguousarray(a_in, dtype=cfg.dtype).reshape(-1))
        r = self.gsum_v(ga, nv, ne, block_size=block_size)
        ga.free()
        return r.get() if isinstance(r, DeviceArray) else r

    def test_sum(self, a_in, N, block_size=256):
        assert 0 <= block_size <= 1024
        ga = DeviceArray.from_host(self.par, np.ascontiguousarray(a_in, dtype=cfg.dtype).reshape(-1))
        r = self.gsum(ga, block_size=block_size)
        ga.free()
        return r
