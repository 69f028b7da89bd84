"""pycuda.gpuarray subset: GPUArray (fill/get/copy/size/dtype/nbytes/gpudata), zeros, empty, *_like, to_gpu, sum."""
import numpy as np
from cuda.bindings import driver as drv

from .driver import DeviceAllocation, check


class GPUArray(object):
    def __init__(self, n, dtype):
        self.dtype = np.dtype(dtype)
        self.size = int(n)
        self.shape = (self.size,)
        self.nbytes = self.size * self.dtype.itemsize
        self.gpudata = DeviceAllocation(self.nbytes)

    def fill(self, v):
        """pyCUDA launches an elementwise fill kernel on the null stream; a driver memset is the same work."""
        if self.nbytes == 0:
            return self
        val = np.asarray(v).astype(self.dtype)
        if not val.tobytes().strip(b"\0"):
            check(drv.cuMemsetD8(int(self.gpudata), 0, self.nbytes))
        elif self.dtype.itemsize == 4:
            check(drv.cuMemsetD32(int(self.gpudata), int(val.view(np.uint32)), self.size))
        else:
            h = np.full(self.size, val, dtype=self.dtype)
            check(drv.cuMemcpyHtoD(int(self.gpudata), h.ctypes.data, h.nbytes))
        return self

    def get(self):
        out = np.empty(self.size, dtype=self.dtype)
        if self.nbytes:
            check(drv.cuMemcpyDtoH(out.ctypes.data, int(self.gpudata), self.nbytes))     # blocking, like pyCUDA
        return out

    def set(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1)
        check(drv.cuMemcpyHtoD(int(self.gpudata), a.ctypes.data, a.nbytes))

    def copy(self):
        g = GPUArray(self.size, self.dtype)
        if self.nbytes:
            check(drv.cuMemcpyDtoD(int(g.gpudata), int(self.gpudata), self.nbytes))
        return g

    def __bool__(self):
        return True

    def __len__(self):
        return self.size


def zeros(n, dtype):
    return GPUArray(n, dtype).fill(0)


def empty(n, dtype):
    return GPUArray(n, dtype)


def empty_like(g):
    return GPUArray(g.size, g.dtype)


def zeros_like(g):
    return GPUArray(g.size, g.dtype).fill(0)


def to_gpu(a):
    a = np.ascontiguousarray(a).reshape(-1)
    g = GPUArray(a.size, a.dtype)
    g.set(a)
    return g


def sum(g):
    """Only reached with use_gpuarray_sum=True (svirl/parallel/reduction.py:57-61), which nothing on the hot path sets."""
    return to_gpu(np.array([g.get().sum()], dtype=g.dtype))
