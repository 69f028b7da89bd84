"""Device context.  Replaces svirl/parallel/startup.py:12-93: instead of JIT-compiling
templated kernel text through pyCUDA, it opens the prebuilt sm_100a library and creates one
``svl_ctx`` for the configured geometry."""
import ctypes as C

import numpy as np

from svirl_b200 import config as cfg
from svirl_b200 import _lib
from .utils import Utils

_active = None


def active():
    """The live Startup (contexts are process-global, like the reference's cfg)."""
    return _active


class Startup(object):

    def __init__(self):
        global _active
        assert isinstance(cfg.device_id, (np.integer, int)) and cfg.device_id >= 0
        solveA = bool(not np.isposinf(cfg.gl_parameter))
        self.reduction_vector_length = 17 if solveA else 5
        self.block_size = 128                       # kept for API parity (reduction tests pass block sizes)
        self.grid_size = Utils.intceil(cfg.N, self.block_size)
        self.grid_size_A = Utils.intceil(cfg.Nab, self.block_size)
        # the reference embeds str(dx) in the kernel text and folds 1/(dx*dx) in double
        dx, dy = float(str(cfg.dx)), float(str(cfg.dy))
        j0, j1 = cfg.slab if cfg.slab is not None else (0, int(cfg.Ny))
        self._ctx = C.c_void_p()
        _lib.call("svl_create", C.byref(self._ctx), int(cfg.device_id), int(cfg.Nx), int(cfg.Ny), dx, dy,
                  int(np.dtype(cfg.dtype).itemsize), int(j0), int(j1))
        self.cuda_compute_capability = (10, 0)
        _active = self

    @property
    def ctx(self):
        if not self._ctx:
            raise _lib.SvirlB200Error("device context already destroyed")
        return self._ctx

    def close(self):
        global _active
        if getattr(self, "_ctx", None):
            _lib.load().svl_destroy(self._ctx)
            self._ctx = C.c_void_p()
        if _active is self:
            _active = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _lib.call("svl_set_option", self.ctx, name.encode(), int(value))

    def stat(self, name):
        v = C.c_double()
        _lib.call("svl_get_stat", self.ctx, name.encode(), C.byref(v))
        return v.value

    def synchronize(self):
        _lib.call("svl_synchronize", self.ctx)

    def get_function(self, function_name):
        """Kernel-level entry points under the reference's kernel names
        (svirl/parallel/startup.py:74-82); see svirl_b200/parallel/kernels.py."""
        from .kernels import lookup
        try:
            return lookup(self, function_name)
        except KeyError:
            raise ValueError("\n Function name %s not found" % (function_name))
