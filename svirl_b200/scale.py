"""TDGL on grids too large for full-size host arrays (new; SURVEY.md row f4 and BASELINE config 5).

``GLSolver`` mirrors every field in a full (Nx, Ny) host array, like the reference.  At 32768^2 and
beyond (65536^2 complex128 = 68 GB) that no longer fits, and with row slabs every rank would hold the
whole grid.  ``ScaleTD`` drives the same library (same kernels, same stop rule, same slab exchange)
through the C ABI with slab-sized host buffers only:

  * the seeded initial order parameter is generated row band by row band from the reference's legacy
    Mersenne-Twister stream (svirl/vars/vars.py:93-109: the first N draws give the moduli, the next N
    the phases, flat index n = i + Nx*j), skipping the draws of other rows, so it is IDENTICAL to
    what ``GLSolver(random_seed=...)`` builds;
  * the symmetric-gauge vector potential (svirl/vars/params.py:135-170) depends on one coordinate
    only and is written row by row;
  * fields are read back by row ranges (``psi_rows``, ``a_rows``, ``b_rows``);
  * ``vortex_count`` runs the GPU winding pass on the slab (the candidate test of
    svirl/observables/vortex_detector.py:62-72 in double precision).
"""
import ctypes as C

import numpy as np

from svirl_b200 import _lib
from svirl_b200.parallel.slab import partition_rows

class _MTStream(object):
    """numpy's legacy Mersenne-Twister double stream (RandomState(seed).random_sample), walked by the library's C
    helper svl_mt19937_doubles: same recurrence on the state numpy seeded, so the values are bit-identical, at ~2 ns
    per draw instead of ~15 (and the GIL is released, so two streams advance concurrently)."""

    def __init__(self, seed, skip=0):
        st = np.random.RandomState(seed).get_state()
        assert st[0] == "MT19937"
        self.key = np.ascontiguousarray(st[1], dtype=np.uint32)
        self.pos = C.c_int(int(st[2]))
        self._pending_skip = int(skip)

    def draw(self, n, out=None):
        if out is None or out.size != int(n):
            out = np.empty(int(n), dtype=np.float64)
        _lib.call("svl_mt19937_doubles", self.key.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(self.pos),
                  self._pending_skip, out.ctypes.data_as(C.POINTER(C.c_double)), int(n))
        self._pending_skip = 0
        return out


class SeededPsi(object):
    """Walks the reference's random initial order parameter row band by row band, starting at row r0.
    psi[n] = (1 - level*u1[n]) * exp(i*pi*level*(2*u2[n] - 1)), u1 = draws [0, N), u2 = draws [N, 2N) of the
    legacy Mersenne Twister seeded with `seed` (np.random.seed + two np.random.rand(N) calls in the reference):
    two streams are positioned at Nx*r0 and N + Nx*r0 once, then advance together."""

    def __init__(self, Nx, Ny, r0, seed, level=1.0, dtype=np.float64):
        self.Nx, self.level = int(Nx), level
        self.ctype = np.complex64 if np.dtype(dtype) == np.float32 else np.complex128
        self.s1 = _MTStream(seed, self.Nx * int(r0))
        self.s2 = _MTStream(seed, self.Nx * int(Ny) + self.Nx * int(r0))
        self._pool = None
        self._free = []          # draw buffers handed back by transform() (first touch of fresh pages is not free)

    def draws(self, nrows):
        """(u1, u2) of the next `nrows` rows; the two streams are walked on two threads."""
        n = self.Nx * int(nrows)
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=2)
        b1 = self._free.pop() if self._free else None
        b2 = self._free.pop() if self._free else None
        f2 = self._pool.submit(self.s2.draw, n, b2)
        u1 = self.s1.draw(n, b1)
        return u1, f2.result()

    def transform(self, u1, u2, nrows):
        """The reference's expression (vars.py:106), numpy ufuncs on the same values -> identical bits."""
        modulus = 1.0 - self.level * u1
        phase = self.level * 1.0j * np.pi * (2.0 * u2 - 1.0)
        out = (modulus * np.exp(phase)).astype(self.ctype).reshape(int(nrows), self.Nx).T      # [i, row]
        self._free.append(u1)
        self._free.append(u2)
        return out

    def transform_rows(self, u1, u2, nrows):
        """Same values as transform() (bit for bit, tests/test_scale_host.py) through the library's host helper
        svl_seeded_psi: no array temporaries, GIL released.  -> C-contiguous (nrows, Nx) array, x fastest."""
        out = np.empty((int(nrows), self.Nx), dtype=self.ctype)
        _lib.call("svl_seeded_psi", u1.ctypes.data_as(C.POINTER(C.c_double)), u2.ctypes.data_as(C.POINTER(C.c_double)),
                  u1.size, float(self.level), out.ctypes.data_as(C.c_void_p), int(np.dtype(self.ctype).itemsize))
        self._free.append(u1)
        self._free.append(u2)
        return out

    def next_rows(self, nrows):
        u1, u2 = self.draws(nrows)
        return self.transform(u1, u2, nrows)


def seeded_psi_rows(Nx, Ny, r0, r1, seed, level=1.0, dtype=np.float64):
    """Rows [r0, r1) of the reference's random initial order parameter, shape (Nx, r1 - r0)."""
    return SeededPsi(Nx, Ny, r0, seed, level, dtype).next_rows(r1 - r0)


def symmetric_gauge_rows(Nx, Ny, dx, dy, H, r0, r1, dtype=np.float64):
    """Rows [r0, r1) of the initial (a, b): a = -(y - Ly/2) H / 2 on a-edges, b = +(x - Lx/2) H / 2 on b-edges
    (edge mid-points), evaluated in `dtype` with the reference's expression order."""
    dt = np.dtype(dtype).type
    Lx, Ly = dt(float(dx) * (Nx - 1)), dt(float(dy) * (Ny - 1))
    dxd, dyd, Hd = dt(dx), dt(dy), dt(H)
    ya = np.linspace(0.0, Ly, num=Ny, endpoint=True, dtype=dt)                       # a-edge rows sit on node rows
    xb = np.linspace(0.0, Lx, num=Nx, endpoint=True, dtype=dt)
    arow = np.zeros(Ny, dtype=dt)
    arow -= 0.5 * (ya - 0.5 * Ly) * Hd
    bcol = np.zeros(Nx, dtype=dt)
    bcol += (1.0 - 0.5) * (xb - 0.5 * Lx) * Hd
    a = np.repeat(arow[None, r0:r1], Nx - 1, axis=0)                                 # (Nx-1, rows)
    rb1 = min(r1, Ny - 1)
    b = np.repeat(bcol[:, None], max(rb1 - r0, 0), axis=1)                           # (Nx, rows of b)
    del dxd, dyd
    return a, b


class _Par(object):
    def __init__(self, ctx):
        self.ctx = ctx


class ScaleTD(object):
    """Slab-local TDGL driver.  One instance per process / GPU; with torch.distributed initialised
    (NCCL) the rows are split over the ranks and neighbours are connected over NVLink."""

    def __init__(self, Nx, Ny, dx=0.5, dy=0.5, dtype=np.float64, gl_parameter=np.inf, normal_conductivity=1.0,
                 homogeneous_external_field=0.0, linear_coefficient=1.0, random_seed=1234, random_level=1.0,
                 device_id=0, distributed=False, band_rows=256):
        self.Nx, self.Ny, self.dx, self.dy = int(Nx), int(Ny), float(dx), float(dy)
        self.dtype = np.dtype(dtype).type
        self.ctype = np.complex64 if self.dtype is np.float32 else np.complex128
        self.kappa, self.sigma, self.H, self.eps = gl_parameter, normal_conductivity, homogeneous_external_field, linear_coefficient
        self.solveA = not np.isposinf(gl_parameter)
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.j0, self.j1 = partition_rows(self.Ny, self.world)[self.rank]
        dxs, dys = float(str(self.dtype(dx))), float(str(self.dtype(dy)))
        self._ctx = C.c_void_p()
        _lib.call("svl_create", C.byref(self._ctx), int(device_id), self.Nx, self.Ny, dxs, dys,
                  int(np.dtype(self.dtype).itemsize), self.j0, self.j1)
        self.par = _Par(self._ctx)
        self.psi, self.ab = C.c_void_p(), C.c_void_p()
        _lib.call("svl_alloc", self._ctx, _lib.NODE_C, 0, 0, C.byref(self.psi))
        _lib.call("svl_alloc", self._ctx, _lib.EDGE, 0, 0, C.byref(self.ab))
        # initial fields, band by band, halo rows included (so no exchange is needed before the first step)
        lo, hi = max(self.j0 - 8, 0), min(self.j1 + 8, self.Ny)
        # the Mersenne-Twister streams are walked sequentially (C helper); the complex exponentials of a band -- the
        # expensive part -- are evaluated on worker threads (svl_seeded_psi, GIL released) while the streams run
        # ahead; uploads happen in band order on this thread
        import os
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        gen = SeededPsi(self.Nx, self.Ny, lo, random_seed, random_level, self.dtype)
        nthreads = max(1, min(32, (os.cpu_count() or 2) - 2))
        pool = ThreadPoolExecutor(max_workers=nthreads)
        pending = deque()

        def _upload(r, r1, fut):
            band = fut.result()
            _lib.call("svl_h2d_rows", self._ctx, self.psi, 0, r, r1, band.ctypes.data_as(C.c_void_p))

        for r in range(lo, hi, band_rows):
            r1 = min(r + band_rows, hi)
            u1, u2 = gen.draws(r1 - r)
            pending.append((r, r1, pool.submit(gen.transform_rows, u1, u2, r1 - r)))
            del u1, u2
            while len(pending) > 2 * nthreads:
                _upload(*pending.popleft())
        while pending:
            _upload(*pending.popleft())
        pool.shutdown()
        for r in range(lo, hi, band_rows):
            r1 = min(r + band_rows, hi)
            a, b = symmetric_gauge_rows(self.Nx, self.Ny, dx, dy, homogeneous_external_field, r, r1, self.dtype)
            fa = np.ascontiguousarray(a.T)
            _lib.call("svl_h2d_rows", self._ctx, self.ab, 0, r, r1, fa.ctypes.data_as(C.c_void_p))
            if b.shape[1] > 0:
                fb = np.ascontiguousarray(b.T)
                _lib.call("svl_h2d_rows", self._ctx, self.ab, 1, r, r + b.shape[1], fb.ctypes.data_as(C.c_void_p))
        del gen
        self.slab_comm = None
        if self.world > 1:
            from svirl_b200.parallel.slab import SlabComm
            self.slab_comm = SlabComm(None, raw=(self.par, self.psi, self.ab, (self.j0, self.j1)))
        self.rand_t = C.c_uint32(int(random_seed) if random_seed is not None else 1)
        self.sweeps = (C.c_longlong * 2)(0, 0)

    # ---- time stepping (svirl/solvers/td.py:342-367 through svl_td_run)
    def td(self, dt, Nt, stop_psi=1e-6, stop_A=1e-6):
        floor = 1e-6 if self.dtype is np.float32 else 1e-12
        k2 = float(self.dtype(self.dtype(self.kappa) ** 2)) if self.solveA else -1.0
        rho = float(self.dtype(1.0 / self.sigma))
        _lib.call("svl_td_run", self._ctx, int(Nt), float(self.dtype(dt)), int(self.solveA), float(self.eps), None, k2, rho,
                  float(self.H), self.psi, self.ab, 0.0, 0.0, C.byref(self.rand_t), float(self.dtype(max(stop_psi, floor))),
                  float(self.dtype(max(stop_A, floor))), self.sweeps)

    def synchronize(self):
        _lib.call("svl_synchronize", self._ctx)

    def stat(self, name):
        v = C.c_double()
        _lib.call("svl_get_stat", self._ctx, name.encode(), C.byref(v))
        return v.value

    # ---- row-range getters (global rows; only rows this rank owns are valid)
    def psi_rows(self, r0, r1):
        out = np.empty((r1 - r0, self.Nx), dtype=self.ctype)
        _lib.call("svl_d2h_rows", self._ctx, out.ctypes.data_as(C.c_void_p), self.psi, 0, int(r0), int(r1))
        return out.T

    def a_rows(self, r0, r1):
        out = np.empty((r1 - r0, self.Nx - 1), dtype=self.dtype)
        _lib.call("svl_d2h_rows", self._ctx, out.ctypes.data_as(C.c_void_p), self.ab, 0, int(r0), int(r1))
        return out.T

    def b_rows(self, r0, r1):
        r1 = min(r1, self.Ny - 1)
        out = np.empty((max(r1 - r0, 0), self.Nx), dtype=self.dtype)
        if r1 > r0:
            _lib.call("svl_d2h_rows", self._ctx, out.ctypes.data_as(C.c_void_p), self.ab, 1, int(r0), int(r1))
        return out.T

    def vortex_candidates(self, cap=1 << 20):
        """(cell indices n = i + (Nx-1) j, winding numbers) of this slab's cells that pass the GPU
        pre-test |v| > 0.45, |v - round v| < 0.15; apply the reference's thresholds with `vortex_count`."""
        while True:
            cells = np.empty(cap, dtype=np.int64)
            vals = np.empty(cap, dtype=np.float64)
            cnt = C.c_size_t()
            _lib.call("svl_vortex_candidates", self._ctx, float(self.H), self.psi, self.ab,
                      cells.ctypes.data_as(C.POINTER(C.c_int64)), vals.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(cnt))
            if cnt.value <= cap:
                k = np.argsort(cells[:cnt.value])
                return cells[:cnt.value][k], vals[:cnt.value][k]
            cap = int(cnt.value)

    def vortex_count(self):
        """Cells of this slab with |v| > 0.5 and |v - round v| < 0.1 (vortex_detector.py:71), split by sign."""
        _, v = self.vortex_candidates()
        ok = (np.abs(v) > 0.5) & (np.abs(v - np.round(v)) < 0.1)
        return int(np.sum(ok & (v > 0))), int(np.sum(ok & (v < 0)))

    def vortices(self, band_rows=512):
        """(x, y, vorticity) of the vortices in this slab's cells, ascending cell index: the GPU candidate pass
        followed by the vectorised host triangulation (observables/triangulate.py, bit-identical to the
        reference's scalar arithmetic on every fixture), row band by row band.  Not yet exercised on hardware
        at full scale; the pieces (candidates, row getters, triangulation) are tested separately."""
        from svirl_b200.observables.triangulate import triangulate
        cells, _ = self.vortex_candidates()
        Nxc = self.Nx - 1
        rows = cells // Nxc
        dx, dy, H = self.dtype(self.dx), self.dtype(self.dy), self.dtype(self.H)
        out = []
        top = min(self.j1, self.Ny - 1)
        for r0 in range(self.j0, top, band_rows):
            r1 = min(r0 + band_rows, top)
            sel = cells[(rows >= r0) & (rows < r1)]
            if sel.size == 0:
                continue
            psi, a, b = self.psi_rows(r0, r1 + 1), self.a_rows(r0, r1 + 1), self.b_rows(r0, r1)
            out.append(triangulate(sel, psi, a, b, a, b, H, dx, dy, np.int32(Nxc), r0, self.dtype))
        if not out:
            z = np.zeros(0, dtype=self.dtype)
            return z, z.copy(), z.copy()
        return tuple(np.concatenate([o[k] for o in out]) for k in range(3))

    # ---- slab checkpoints: one .npz per rank holding the rows this rank owns
    def save_checkpoint(self, prefix):
        """Write <prefix>.rank<r>of<w>.npz: the owned rows of psi, a, b (host layout [i, row]), the Langevin step counter
        and the sweep counters.  Restoring it on the same decomposition continues the run bit for bit."""
        self.synchronize()
        j0, j1 = self.j0, self.j1
        path = "%s.rank%dof%d.npz" % (prefix, self.rank, self.world)
        np.savez(path, psi=self.psi_rows(j0, j1), a=self.a_rows(j0, j1), b=self.b_rows(j0, j1),
                 rows=np.array([j0, j1]), shape=np.array([self.Nx, self.Ny]), rand_t=np.uint32(self.rand_t.value),
                 sweeps=np.array([self.sweeps[0], self.sweeps[1]], dtype=np.int64),
                 params=np.array([self.dx, self.dy, float(self.kappa), float(self.sigma), float(self.H), float(self.eps)]))
        return path

    def load_checkpoint(self, prefix):
        """Inverse of save_checkpoint (same grid, same number of ranks); refreshes the halo rows from the neighbours."""
        path = "%s.rank%dof%d.npz" % (prefix, self.rank, self.world)
        with np.load(path) as d:
            assert tuple(d["shape"]) == (self.Nx, self.Ny) and tuple(d["rows"]) == (self.j0, self.j1), "checkpoint of another decomposition"
            j0, j1 = self.j0, self.j1
            psi = np.ascontiguousarray(d["psi"].T.astype(self.ctype))
            _lib.call("svl_h2d_rows", self._ctx, self.psi, 0, j0, j1, psi.ctypes.data_as(C.c_void_p))
            a = np.ascontiguousarray(d["a"].T.astype(self.dtype))
            _lib.call("svl_h2d_rows", self._ctx, self.ab, 0, j0, j1, a.ctypes.data_as(C.c_void_p))
            b = np.ascontiguousarray(d["b"].T.astype(self.dtype))
            if b.shape[0] > 0:
                _lib.call("svl_h2d_rows", self._ctx, self.ab, 1, j0, j0 + b.shape[0], b.ctypes.data_as(C.c_void_p))
            self.rand_t = C.c_uint32(int(d["rand_t"]))
            self.sweeps[0], self.sweeps[1] = int(d["sweeps"][0]), int(d["sweeps"][1])
        if self.world > 1:
            _lib.call("svl_slab_exchange", self._ctx, self.psi)
            _lib.call("svl_slab_exchange", self._ctx, self.ab)
        _lib.call("svl_set_option", self._ctx, b"reset_prediction", 0)

    def close(self):
        if getattr(self, "_ctx", None):
            lib = _lib.load()
            lib.svl_free(self._ctx, self.psi)
            lib.svl_free(self._ctx, self.ab)
            lib.svl_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
