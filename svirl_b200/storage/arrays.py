"""Host/device paired arrays over the C ABI.

``GArray`` keeps the public surface of svirl/storage/arrays.py:11-609 (constructor keywords,
``sync``/``need_htod_sync``/``need_dtoh_sync``/``synced``, ``get_h``/``get_d``/``get_d_obj``/
``get_vec_h``/``set_h``/``set_vec_h``, ``free``, ``metadata``); ``DeviceArray`` plays the role
pyCUDA's ``GPUArray`` plays there (``get``, ``fill``, ``copy``, ``size``, ``gpudata.free()``).

Host arrays are indexed [i, j]; the flat layout is x-fastest: flat = reshape(arr.T, n)
(arrays.py:327-350).  On the device, fields whose shape matches the solver geometry live in
pitched planes (node / edge / cell kinds); anything else is a flat array.
"""
import ctypes as C
from warnings import warn

import numpy as np

from svirl_b200 import config as cfg
from svirl_b200 import _lib


def _par():
    from svirl_b200.parallel import startup
    p = startup.active()
    if p is None:
        raise _lib.SvirlB200Error("no active device context (construct GLSolver first)")
    return p


def _kind_for(shape, dtype):
    """Device kind for a host shape/dtype under the current geometry."""
    dtype = np.dtype(dtype)
    if cfg.Nx is None:
        return _lib.FLAT
    Nx, Ny = int(cfg.Nx), int(cfg.Ny)
    if isinstance(shape, list):
        if len(shape) == 2 and tuple(shape[0]) == (Nx - 1, Ny) and tuple(shape[1]) == (Nx, Ny - 1) \
                and dtype.kind == 'f':
            return _lib.EDGE
        return _lib.FLAT
    shape = tuple(int(s) for s in shape)
    if shape == (Nx, Ny):
        if dtype.kind == 'c':
            return _lib.NODE_C
        if dtype.kind == 'f':
            return _lib.NODE_R
    if shape == (Nx - 1, Ny - 1):
        if dtype.kind == 'b':
            return _lib.CELL_B
        if dtype.kind == 'f':
            return _lib.CELL_R
    return _lib.FLAT


def _device_dtype(kind, dtype):
    if kind == _lib.NODE_C:
        return np.dtype(cfg.dtype_complex)
    if kind in (_lib.NODE_R, _lib.EDGE, _lib.CELL_R):
        return np.dtype(cfg.dtype)
    if kind == _lib.CELL_B:
        return np.dtype(np.bool_)
    return np.dtype(dtype)


class _Alloc(object):
    """Stands in for GPUArray.gpudata: only ``free()`` is used by callers."""

    def __init__(self, owner):
        self._owner = owner

    def free(self):
        self._owner.free()


class DeviceArray(object):
    """A device buffer handle (``svl_buf*``) with the slice of pyCUDA GPUArray's API that the
    solver layer and the reference's tests use."""

    def __init__(self, par, kind, size, dtype):
        self.par = par
        self.kind = kind
        self.dtype = np.dtype(dtype)
        self.handle = C.c_void_p()
        _lib.call("svl_alloc", par.ctx, kind, int(size), int(self.dtype.itemsize), C.byref(self.handle))
        self.size = int(_lib.load().svl_buf_size(self.handle))
        self.gpudata = _Alloc(self)

    @classmethod
    def from_host(cls, par, flat, kind=_lib.FLAT):
        flat = np.ascontiguousarray(flat).reshape(-1)
        d = cls(par, kind, flat.size, flat.dtype)
        d.set(flat)
        return d

    @classmethod
    def zeros(cls, par, kind, size=0, dtype=None):
        return cls(par, kind, size, _device_dtype(kind, dtype if dtype is not None else cfg.dtype))

    @property
    def shape(self):
        return (self.size,)

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    def _check(self):
        if not self.handle:
            raise _lib.SvirlB200Error("device array already freed")

    def get(self):
        """Flat host copy in the reference's device layout."""
        self._check()
        # zeros, not empty: on a slab context svl_d2h fills the owned rows only, and a later push() must not
        # upload uninitialised memory into the halo rows (only GLSolver.owned_rows() are valid on slabs)
        out = np.zeros(self.size, dtype=self.dtype)
        _lib.call("svl_d2h", self.par.ctx, out.ctypes.data_as(C.c_void_p), self.handle)
        return out

    def set(self, flat):
        self._check()
        flat = np.ascontiguousarray(flat, dtype=self.dtype).reshape(-1)
        assert flat.size == self.size, (flat.size, self.size)
        _lib.call("svl_h2d", self.par.ctx, self.handle, flat.ctypes.data_as(C.c_void_p))

    def fill(self, value):
        self._check()
        if value == 0:
            _lib.call("svl_fill_zero", self.par.ctx, self.handle)
        else:
            self.set(np.full(self.size, value, dtype=self.dtype))
        return self

    def copy(self):
        d = DeviceArray(self.par, self.kind, self.size, self.dtype)
        d.copy_from(self)
        return d

    def copy_from(self, src):
        self._check()
        _lib.call("svl_d2d", self.par.ctx, self.handle, src.handle)

    def swap(self, other):
        _lib.call("svl_swap", self.par.ctx, self.handle, other.handle)

    def free(self):
        if getattr(self, "handle", None) and getattr(self.par, "_ctx", None):
            _lib.load().svl_free(self.par._ctx, self.handle)
        self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __bool__(self):
        return True


class GArray(object):
    """Host + device paired array with lazy synchronisation flags."""

    on_device = 'd'
    on_host_and_device = 'hd'

    def __init__(self, nelements=None, dtype=np.float64, shape=None, on=None, like=None,
                 like_init=True, name='array-0', suppress_warnings=False):
        self._quiet = suppress_warnings
        if on is not None and on not in (self.on_device, self.on_host_and_device):
            raise ValueError("`on` must be GArray.on_device or GArray.on_host_and_device")
        if like is not None and not isinstance(like, (np.ndarray, DeviceArray, GArray)):
            raise TypeError("`like` accepts numpy arrays, DeviceArray or GArray")
        if like is None and nelements is None and shape is None:
            raise ValueError("give `nelements`, `shape` or `like`")
        if like is None and np.dtype(dtype).type not in (np.float32, np.float64, np.complex64, np.complex128):
            raise TypeError("`dtype` must be a float or complex type")
        if like is not None and shape is not None:
            self._warn("shape ignored because a prototype is given")
        if like is None and nelements is not None and shape is not None:
            self._warn("both nelements and shape given; using shape")
            nelements = None

        self.name = str(name)
        self._data = None          # flat host mirror
        self._data_v = None        # per-component views for vector storage
        self._gdata = None         # DeviceArray
        self._ndim = 1
        self._nelements_v = None
        self.__sync_status = None  # -1 host newer, 0 equal, +1 device newer
        self._on = on or self.on_host_and_device

        init = None
        if like is not None:
            self._shape = like.shape
            self._dtype = like.dtype
            self._nelements = int(like.size)
            if isinstance(like, GArray) and like._ndim > 1:
                self._ndim = like._ndim
                self._nelements_v = like._nelements_v.copy()
            if like_init:
                if isinstance(like, np.ndarray):
                    init = ('host', self._flatten(like))
                elif isinstance(like, GArray):
                    init = ('dev', like.get_d_obj())
                else:
                    init = ('dev', like)
        else:
            self._dtype = np.dtype(dtype)
            if shape is not None:
                self._shape = shape
                if isinstance(shape, list):
                    self._ndim = len(shape)
                    assert 0 < self._ndim < 3
                    sizes = np.array([int(np.prod(s)) for s in shape], dtype=np.int64)
                    self._nelements = int(sizes.sum())
                    self._nelements_v = np.cumsum(sizes)
                else:
                    self._nelements = int(np.prod(shape))
            else:
                self._shape = (int(nelements), 1)
                self._nelements = int(nelements)

        par = _par()
        kind = _kind_for(self._shape, self._dtype)
        if isinstance(like, DeviceArray):
            kind = like.kind
        elif isinstance(like, GArray) and like._gdata is not None:
            kind = like._gdata.kind
        ddtype = _device_dtype(kind, self._dtype)
        self._dtype = ddtype if kind != _lib.FLAT else np.dtype(self._dtype)
        self._gdata = DeviceArray(par, kind, self._nelements, self._dtype)
        if self._on == self.on_host_and_device:
            self._data = np.zeros(self._nelements, dtype=self._dtype)
            self.synced()
        if init is not None:
            if init[0] == 'host':
                flat = init[1].astype(self._dtype, copy=False)
                self._gdata.set(flat)
                if self._data is not None:
                    np.copyto(self._data, flat)
            else:
                self._gdata.copy_from(init[1])
                if self._data is not None:
                    np.copyto(self._data, self._gdata.get())
        if self._ndim > 1 and self._data is not None:
            self._data_v = np.split(self._data, self._nelements_v)

    # ---- properties
    @property
    def size(self):
        return self._nelements

    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        return self._shape

    @property
    def on(self):
        return self._on

    # ---- synchronisation
    def _hd(self):
        return self._on == self.on_host_and_device

    def sync(self):
        if not self._hd():
            return
        if self.__sync_status is not None and self.__sync_status > 0:
            np.copyto(self._data, self._gdata.get())
        elif self.__sync_status is not None and self.__sync_status < 0:
            self._gdata.set(self._data)
        self.synced()

    def push(self):
        """Make the device copy current WITHOUT pulling device data back: host-to-device copy
        only if the host side was modified (used at solver entry; the reference's solvers do not
        read the fields back between calls either)."""
        if self._hd() and self.__sync_status is not None and self.__sync_status < 0:
            self._gdata.set(self._data)
            self.synced()

    def need_dtoh_sync(self):
        if self._hd():
            self.__sync_status = 1

    def need_htod_sync(self):
        if self._hd():
            self.__sync_status = -1

    def synced(self):
        if self._hd():
            self.__sync_status = 0

    # ---- layout helpers: host [i, j]  <->  flat x-fastest
    @staticmethod
    def _flatten_to(arr, size):
        return None if arr is None else np.reshape(arr.T, size)

    def _flatten(self, arr):
        return self._flatten_to(arr, self._nelements)

    @staticmethod
    def _unflatten(flat, shape):
        return None if flat is None else np.reshape(flat, tuple(reversed(shape))).T

    # ---- getters (views of the host mirror, not copies -- callers edit them in place)
    def get_h(self, sync=True):
        if sync:
            self.sync()
        if self._ndim > 1:
            return None
        return self._unflatten(self._data, self._shape)

    def get_d(self, sync=True):
        if sync:
            self.sync()
        if self._ndim > 1:
            return None
        return self._unflatten(self._gdata.get(), self._shape)

    def get_d_obj(self, sync=False):
        if sync:
            self.sync()
        return self._gdata

    def get_vec_h(self, sync=True):
        if sync:
            self.sync()
        if self._ndim > 1:
            return (self._unflatten(self._data_v[0], self._shape[0]),
                    self._unflatten(self._data_v[1], self._shape[1]))
        return (None, None)

    # ---- setters
    def set_h(self, arr):
        if arr.size != self._nelements:
            warn("Size mismatch: Storage has size %d but input array has: %d" % (self._nelements, arr.size))
            return
        np.copyto(self._data, self._flatten(arr), casting='unsafe')
        self.need_htod_sync()

    def set_vec_h(self, arr_a, arr_b):
        assert arr_a.shape == tuple(self._shape[0])
        assert arr_b.shape == tuple(self._shape[1])
        if self._data is not None:
            np.copyto(self._data_v[0], self._flatten_to(arr_a, arr_a.size), casting='unsafe')
            np.copyto(self._data_v[1], self._flatten_to(arr_b, arr_b.size), casting='unsafe')
            self.need_htod_sync()

    def free(self):
        if self._gdata is not None:
            self._gdata.free()
            self._gdata = None
        self._data = None
        self._data_v = None
        self._nelements = 0
        self._shape = None
        self._dtype = None

    def metadata(self):
        for k, v in (('Name', self.name), ('size', self._nelements), ('shape', self._shape),
                     ('dtype', self._dtype), ('on', self._on), ('sync status', self.__sync_status)):
            print('  %-11s: ' % k, v, flush=True)
        print('', flush=True)

    def _warn(self, msg):
        if not self._quiet:
            warn(msg)
