import numpy as np


class _DevAlloc(object):
    def __init__(self, owner):
        self.owner = owner

    def free(self):
        pass


class GPUArray(object):
    def __init__(self, arr):
        self._a = np.ascontiguousarray(arr).reshape(-1) if np.ndim(arr) else np.array(arr).reshape(())
        self.gpudata = _DevAlloc(self)

    @property
    def size(self):
        return self._a.size

    @property
    def dtype(self):
        return self._a.dtype

    @property
    def shape(self):
        return self._a.shape

    @property
    def nbytes(self):
        return self._a.nbytes

    def get(self):
        return self._a.copy()

    def fill(self, v):
        self._a[...] = v
        return self

    def copy(self):
        return GPUArray(self._a.copy())

    def __bool__(self):
        return True

    def __len__(self):
        return self._a.size


def zeros(n, dtype):
    return GPUArray(np.zeros(int(n), dtype=dtype))


def empty(n, dtype):
    return GPUArray(np.zeros(int(n), dtype=dtype))


def empty_like(g):
    return GPUArray(np.zeros(g.size, dtype=g.dtype))


def zeros_like(g):
    return GPUArray(np.zeros(g.size, dtype=g.dtype))


def to_gpu(a):
    return GPUArray(np.array(a).reshape(-1).copy())


def sum(g):
    return GPUArray(np.array(g._a.sum()))
