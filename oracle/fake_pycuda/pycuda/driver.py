import numpy as np


def init():
    pass


class _Context(object):
    def pop(self):
        pass


class Device(object):
    def __init__(self, dev_id):
        self.id = dev_id

    def compute_capability(self):
        return (10, 0)

    def make_context(self):
        return _Context()


def _buf(x):
    return x.owner._a if hasattr(x, "owner") else x


def memcpy_dtod(dest, src, nbytes):
    d, s = _buf(dest), _buf(src)
    dv = d.reshape(-1).view(np.uint8)
    sv = s.reshape(-1).view(np.uint8)
    dv[:nbytes] = sv[:nbytes]


def memcpy_htod(dest, src):
    d = _buf(dest)
    d.reshape(-1).view(np.uint8)[:src.nbytes] = np.ascontiguousarray(src).reshape(-1).view(np.uint8)


def memcpy_dtoh(dest, src):
    s = _buf(src)
    dest.reshape(-1).view(np.uint8)[:] = s.reshape(-1).view(np.uint8)[:dest.nbytes]
