"""Row-slab decomposition across GPUs (new; the reference is single-GPU).

One process per GPU (torchrun); rank r owns node rows [j0, j1) of the global grid.  The
library moves halo rows by direct peer stores over NVLink (CUDA IPC handles exchanged here
through torch.distributed) and calls back into `_reduce_max` for the only collective of the
TDGL path: the MAX of the per-sweep residual slots.  torch.distributed is plumbing only."""
import ctypes as C

import numpy as np

from svirl_b200 import _lib

HANDLE_BYTES = 64 + 8 * 10    # struct SlabHandle in slab.cu: IPC handle of the arena + 10 offsets


def partition_rows(Ny, world):
    """Balanced contiguous row ranges: [(j0, j1)] * world, sizes differ by at most one."""
    base, extra = divmod(int(Ny), int(world))
    out, j = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((j, j + n))
        j += n
    return out


def max_reduce_u64(values, group=None):
    """In-place MAX over ranks of a numpy uint64 array (bit patterns of non-negative doubles order
    like integers, so this is the exact max of the residuals).  Works with gloo (CPU tensors) and
    nccl (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(values.view(np.int64))
    if dist.get_backend(group) == "nccl":
        d = t.cuda()
        dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
        t.copy_(d.cpu())
    else:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return values


class SlabComm(object):
    """Connects this rank's context to its neighbours and installs the residual reduction."""

    def __init__(self, gl, group=None, raw=None):
        """gl: a GLSolver built with slab=...; or raw = (par_like_with_ctx, psi_handle, ab_handle, (j0, j1)) for
        drivers that own their buffers (svirl_b200/scale.py)."""
        import torch.distributed as dist
        self.gl, self.group = gl, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if raw is None:
            raw = (gl.par, gl.vars.order_parameter_h().handle, gl.vars.vector_potential_h().handle, gl.cfg.slab)
        par, psi_handle, ab_handle, slab = raw
        self._par, self._slab = par, tuple(int(x) for x in slab)
        mine = (C.c_char * HANDLE_BYTES)()
        _lib.call("svl_slab_export", par.ctx, psi_handle, ab_handle, C.cast(mine, C.c_void_p))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (bytes(mine), self._slab), group=group)
        lo = gathered[self.rank - 1] if self.rank > 0 else None
        hi = gathered[self.rank + 1] if self.rank + 1 < self.world else None
        self._keep = [C.create_string_buffer(x[0], HANDLE_BYTES) if x else None for x in (lo, hi)]
        _lib.call("svl_slab_connect", par.ctx,
                  C.cast(self._keep[0], C.c_void_p) if lo else None, lo[1][0] if lo else 0,
                  C.cast(self._keep[1], C.c_void_p) if hi else None, hi[1][0] if hi else 0)

        # residual board: all-to-all peer memory for the MAX of the residual slots (NVLink, no NCCL)
        if dist.get_backend(group) == "nccl" and self.world <= 16:
            bh = (C.c_char * 64)()
            _lib.call("svl_slab_board_export", par.ctx, C.cast(bh, C.c_void_p))
            allh = [None] * self.world
            dist.all_gather_object(allh, bytes(bh), group=group)
            self._board_handles = C.create_string_buffer(b"".join(allh), 64 * self.world)
            _lib.call("svl_slab_board_connect", par.ctx, self.rank, self.world, C.cast(self._board_handles, C.c_void_p))

        @C.CFUNCTYPE(None, C.POINTER(C.c_ulonglong), C.c_int)
        def _reduce(ptr, n):
            arr = np.ctypeslib.as_array(ptr, shape=(n,))
            max_reduce_u64(arr, group)

        self._cb = _reduce                                   # keep the callback alive
        _lib.call("svl_set_reduce_callback", par.ctx, C.cast(self._cb, C.c_void_p))
        if dist.get_backend(group) == "nccl":
            # NCCL can reduce the device words in place, ordered on the library's own stream
            import torch
            stream = torch.cuda.ExternalStream(int(_lib.load().svl_get_stream(par.ctx)))

            class _Dev(object):
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

            @C.CFUNCTYPE(None, C.c_void_p, C.c_int)
            def _reduce_dev(ptr, n):
                with torch.cuda.stream(stream):
                    t = torch.as_tensor(_Dev(int(ptr), int(n)), device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)

            self._cb_dev = _reduce_dev
            _lib.call("svl_set_reduce_callback_device", par.ctx, C.cast(self._cb_dev, C.c_void_p))
        dist.barrier(group=group)

    def exchange(self, garray):
        """Refresh the halo rows of psi / A from the neighbours (after host-side edits)."""
        garray.push()
        _lib.call("svl_slab_exchange", self._par.ctx, garray.get_d_obj().handle)

    def owned_rows(self, arr2d):
        """Rows of a host array indexed [i, j] that this rank owns (node rows)."""
        j0, j1 = self._slab
        return arr2d[:, j0:min(j1, arr2d.shape[1])]
