"""pycuda.driver subset: init, Device.compute_capability/make_context, memcpy_{dtod,htod,dtoh}."""
import ctypes

import numpy as np
from cuda.bindings import driver as drv

_state = {"init": False, "ctx": {}}


def check(res):
    """cuda-python returns (CUresult, value...) tuples."""
    err = res[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        _, name = drv.cuGetErrorName(err)
        raise RuntimeError("CUDA driver error: %s" % (name.decode() if name else err))
    return res[1] if len(res) == 2 else res[1:]


def init():
    if not _state["init"]:
        check(drv.cuInit(0))
        _state["init"] = True


class _Context(object):
    """The device's PRIMARY context (shared with the CUDA runtime, so libsvirl_b200 can live in the same
    process for three-way parity tests); pyCUDA would create a new one, which changes nothing numerically."""

    def __init__(self, dev):
        self.dev = dev
        self.handle = check(drv.cuDevicePrimaryCtxRetain(dev))
        check(drv.cuCtxSetCurrent(self.handle))

    def pop(self):
        pass

    def synchronize(self):
        check(drv.cuCtxSynchronize())


class Device(object):
    def __init__(self, dev_id):
        init()
        self.id = int(dev_id)
        self.handle = check(drv.cuDeviceGet(self.id))

    def compute_capability(self):
        major = check(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, self.handle))
        minor = check(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, self.handle))
        return (major, minor)

    def make_context(self):
        if self.id not in _state["ctx"]:
            _state["ctx"][self.id] = _Context(self.handle)
        else:
            check(drv.cuCtxSetCurrent(_state["ctx"][self.id].handle))
        return _state["ctx"][self.id]


def synchronize():
    check(drv.cuCtxSynchronize())


class DeviceAllocation(object):
    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = int(check(drv.cuMemAlloc(max(self.nbytes, 1))))

    def free(self):
        if self.ptr:
            drv.cuMemFree(self.ptr)
            self.ptr = 0

    def __int__(self):
        return self.ptr

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(x):
    return int(x.gpudata) if hasattr(x, "gpudata") else int(x)


def memcpy_dtod(dest, src, nbytes):
    check(drv.cuMemcpyDtoD(_ptr(dest), _ptr(src), int(nbytes)))


def memcpy_htod(dest, src):
    src = np.ascontiguousarray(src)
    check(drv.cuMemcpyHtoD(_ptr(dest), src.ctypes.data, src.nbytes))


def memcpy_dtoh(dest, src):
    assert dest.flags["C_CONTIGUOUS"]
    check(drv.cuMemcpyDtoH(dest.ctypes.data, _ptr(src), dest.nbytes))
