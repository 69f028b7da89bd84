"""NumPy oracle for svirl's TDGL / CG hot path -- TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (vectorised NumPy) of the discretisation that
microsoft/svirl implements in its pyCUDA kernels and host loops.  It is the
*checker* for the CUDA path in ``svirl_b200``; it is never the product:

    only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity status: PINNED.  ``oracle/build_ref.py`` compiles the reference's own
kernel sources (where they lie under /root/reference) for the CPU behind a SIMT
shim, ``oracle/refrun.py`` runs the unmodified reference Python package on top
of it, and ``tests/golden/*.npz`` (made by ``oracle/make_golden.py``) hold the
reference's outputs; ``tests/test_oracle_golden.py`` checks every function here
against them.

Array conventions (same as the reference's public API, svirl/storage/arrays.py:
327-350): host arrays are indexed ``[i, j]`` (x first), shapes
psi (Nx,Ny) complex, a (Nx-1,Ny), b (Nx,Ny-1), mt (Nx-1,Ny-1) bool, cells
(Nx-1,Ny-1).  ``U(phi) = exp(-i phi)`` (svirl/cuda/common.h:29-33).

Each function cites the reference file:line it follows.
"""

import numpy as np

__all__ = [
    "Grid", "node_flags", "psi_sweep", "a_sweep", "stop_test", "td_psi_solve", "td_a_solve",
    "td_run", "free_energy", "jacobian_psi", "jacobian_A", "coef_psi", "coef", "alpha_psi_min",
    "alpha_min", "cg_run", "magnetic_field", "current_density", "supercurrent_density",
    "vortices", "initial_psi", "initial_A", "rand_hash", "rand_1", "rand_2",
]


class Grid(object):
    """Geometry + dtype bundle (svirl/__init__.py:67-117)."""

    def __init__(self, Nx, Ny, dx, dy, dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.ctype = {np.float32: np.complex64, np.float64: np.complex128}[self.dtype]
        self.Nx, self.Ny = int(Nx), int(Ny)
        self.dx, self.dy = self.dtype(dx), self.dtype(dy)
        # The kernels embed str(dx) and fold 1/(dx*dx) in DOUBLE even for fp32
        # (svirl/parallel/startup.py:54-60, svirl/cuda/td.h:34-35).
        dxd, dyd = float(str(self.dx)), float(str(self.dy))
        r = self.dtype
        self.idx, self.idy = r(1.0 / dxd), r(1.0 / dyd)
        self.idx2, self.idy2 = r(1.0 / (dxd * dxd)), r(1.0 / (dyd * dyd))
        self.idxy = r(1.0 / (dxd * dyd))
        self.Lx, self.Ly = self.dtype(self.dx * (self.Nx - 1)), self.dtype(self.dy * (self.Ny - 1))


# ----------------------------------------------------------------------------------------------
# material flags
# ----------------------------------------------------------------------------------------------

def node_flags(g, mt):
    """mm, mp, pm, pp per node (svirl/cuda/td.h:48-57): the four cells around
    node (i,j); out-of-range cells are False; no tiling => all in-range True."""
    Nx, Ny = g.Nx, g.Ny
    P = np.zeros((Nx + 1, Ny + 1), dtype=bool)
    P[1:-1, 1:-1] = True if mt is None else np.asarray(mt, dtype=bool)
    return P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]


def _shift(arr, di, dj):
    """out[i,j] = arr[i+di, j+dj], zero outside."""
    out = np.zeros_like(arr)
    Nx, Ny = arr.shape
    si = slice(max(0, -di), min(Nx, Nx - di))
    sj = slice(max(0, -dj), min(Ny, Ny - dj))
    so_i = slice(si.start + di, si.stop + di)
    so_j = slice(sj.start + dj, sj.stop + dj)
    out[si, sj] = arr[so_i, so_j]
    return out


def _pad_a(g, a):
    """a on (Nx-1,Ny) -> (Nx,Ny) with zero column i=Nx-1."""
    out = np.zeros((g.Nx, g.Ny), dtype=a.dtype)
    out[:-1, :] = a
    return out


def _pad_b(g, b):
    out = np.zeros((g.Nx, g.Ny), dtype=b.dtype)
    out[:, :-1] = b
    return out


def _U(ph):
    """exp(-i ph) = cos - i sin (svirl/cuda/common.h:29-33)."""
    return np.cos(ph) - 1j * np.sin(ph)


# ----------------------------------------------------------------------------------------------
# Langevin hash RNG (svirl/cuda/common.h:36-63)
# ----------------------------------------------------------------------------------------------

def rand_hash(s):
    s = np.asarray(s, dtype=np.uint32)
    with np.errstate(over="ignore"):
        s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
        s = s * np.uint32(9)
        s = s ^ (s >> np.uint32(4))
        s = s * np.uint32(0x27d4eb2d)
        s = s ^ (s >> np.uint32(15))
    return s


def rand_1(n, t, dtype):
    with np.errstate(over="ignore"):
        s = np.uint32(71) * np.asarray(n, dtype=np.uint32) + np.uint32(9887) * np.uint32(t)
    # double literal * real_t(hash): product in double, then rounded to real_t
    return (2.0 ** -32 * rand_hash(s).astype(dtype).astype(np.float64)).astype(dtype)


def rand_2(n, t, dtype):
    with np.errstate(over="ignore"):
        s = np.uint32(73) * np.asarray(n, dtype=np.uint32) + np.uint32(9901) * np.uint32(t) + np.uint32(1)
    return (2.0 ** -32 * rand_hash(s).astype(dtype).astype(np.float64)).astype(dtype)


# ----------------------------------------------------------------------------------------------
# E1: one Jacobi sweep of the backward-Euler psi equation (svirl/cuda/td.h:5-133)
# ----------------------------------------------------------------------------------------------

def psi_sweep(g, dt, eps, mt, a, b, psi_rhs, psi, langevin_c=0.0, jstep=1, rand_t=1):
    """Returns (psi_next, r, psi_rhs) with r = max over ALL nodes of
    max(|dRe|,|dIm|) (td.h:116-132, quirk Q8).  WA edge weights (td.h:78-82).
    a, b are the regular(+irregular) potential, not the external one.
    psi_rhs is returned because sweep 0 writes the Langevin noise back (td.h:66-70)."""
    r_t, c_t = g.dtype, g.ctype
    dt = r_t(dt)
    mm, mp, pm, pp = node_flags(g, mt)
    active = mm | mp | pm | pp
    wW, wE, wS, wN = (mm | mp), (pm | pp), (mm | pm), (mp | pp)
    if jstep == 0 and langevin_c > 1.0e-32:
        n = (np.arange(g.Nx)[:, None] + g.Nx * np.arange(g.Ny)[None, :]).astype(np.uint32)
        noise = (rand_1(n, rand_t, r_t) - r_t(0.5)) + 1j * (rand_2(n, rand_t, r_t) - r_t(0.5))
        psi_rhs = np.where(active, psi_rhs + r_t(langevin_c) * noise.astype(c_t), psi_rhs).astype(c_t)
    ap, bp = _pad_a(g, a), _pad_b(g, b)
    tW = wW * (_U(-g.dx * _shift(ap, -1, 0)).astype(c_t) * _shift(psi, -1, 0))
    tE = wE * (_U(g.dx * ap).astype(c_t) * _shift(psi, 1, 0))
    tS = wS * (_U(-g.dy * _shift(bp, 0, -1)).astype(c_t) * _shift(psi, 0, -1))
    tN = wN * (_U(g.dy * bp).astype(c_t) * _shift(psi, 0, 1))
    eps_n = np.asarray(eps, dtype=r_t)
    diag = r_t(1.0) + dt * (psi_rhs.real ** 2 + psi_rhs.imag ** 2 - eps_n
                            + (g.idx2 * (wW.astype(r_t) + wE) + g.idy2 * (wS.astype(r_t) + wN)))
    nxt = (psi_rhs + dt * (g.idx2 * (tW + tE) + g.idy2 * (tS + tN))) / diag
    nxt = np.where(active, nxt, 0).astype(c_t)
    d = nxt - psi
    r = max(np.abs(d.real).max(), np.abs(d.imag).max())
    return nxt, r_t(r), psi_rhs


def stop_test(r, stop_eps, dtype):
    """Exact reference stop decision (td.h:124-132 + td.py:198-201):
    int32(real_t(1e4*r/eps) clamped at 1e8) -> 1e-4*int < 1  <=>  int < 10000."""
    v = dtype(1.0e4 * float(r) / float(dtype(stop_eps)))
    if v > 1.0e8:
        v = dtype(1.0e8)
    return int(v) < 10000


def td_psi_solve(g, dt, eps, mt, a, b, psi, stop_eps=1e-6, langevin_c=0.0, rand_t=1, max_sweeps=1024):
    """psi-solve driver (svirl/solvers/td.py:157-218). Returns (psi_new, sweeps)."""
    rhs = psi.copy()
    cur = psi
    for j in range(max_sweeps):
        cur, r, rhs = psi_sweep(g, dt, eps, mt, a, b, rhs, cur, langevin_c, j, rand_t)
        if stop_test(r, stop_eps, g.dtype):
            break
    return cur, j + 1


# ----------------------------------------------------------------------------------------------
# E2: one Jacobi sweep of the vector-potential equation (svirl/cuda/td.h:311-463)
# ----------------------------------------------------------------------------------------------

def _js(psi0, ph, psi1):
    """Im(conj(psi0) U(ph) psi1) (svirl/cuda/common.h:65-73)."""
    x0, y0, x1, y1 = psi0.real, psi0.imag, psi1.real, psi1.imag
    return (x0 * y1 - y0 * x1) * np.cos(ph) - (x0 * x1 + y0 * y1) * np.sin(ph)


def a_sweep(g, dt, kappa2, rho, H, mt, psi, a_ph, b_ph, a_rhs, b_rhs, a, b,
            langevin_c=0.0, jstep=1, rand_t=1):
    """Returns (a_next, b_next, r, a_rhs, b_rhs).  (a_ph, b_ph) is the kernel's
    ``abi_ab_rhs`` argument (td.h:319, quirk Q1: the caller decides which iterate)."""
    r_t = g.dtype
    Nx, Ny = g.Nx, g.Ny
    dt, kappa2, rho, H = r_t(dt), r_t(kappa2), r_t(rho), r_t(H)
    dt_rho = dt * rho
    dtrk = dt_rho * kappa2
    mm, mp, pm, pp = node_flags(g, mt)
    Na = (Nx - 1) * Ny
    if jstep == 0 and langevin_c > 1.0e-32:
        na = (np.arange(Nx - 1)[:, None] + (Nx - 1) * np.arange(Ny)[None, :]).astype(np.uint32)
        nb = (Na + np.arange(Nx)[:, None] + Nx * np.arange(Ny - 1)[None, :]).astype(np.uint32)
        a_rhs = (a_rhs + r_t(langevin_c) * (rand_1(na, rand_t, r_t) - r_t(0.5))).astype(r_t)
        b_rhs = (b_rhs + r_t(langevin_c) * (rand_2(nb, rand_t, r_t) - r_t(0.5))).astype(r_t)

    # ---- a edges (i < Nx-1, all j) : td.h:366-410
    rh = np.zeros((Nx - 1, Ny), dtype=r_t)
    dd = np.ones((Nx - 1, Ny), dtype=r_t)
    rh[:, 0] = r_t(2.0) * kappa2 * H * g.idy
    dd[:, 0] = 2
    rh[:, Ny - 1] = -r_t(2.0) * kappa2 * H * g.idy
    dd[:, Ny - 1] = 2
    on = (pm | pp)[:-1, :]
    jl = np.where(on, g.idx * _js(psi[:-1, :], g.dx * a_ph, psi[1:, :]), 0).astype(r_t)
    lo = np.zeros((Nx - 1, Ny), dtype=r_t)
    hi = np.zeros((Nx - 1, Ny), dtype=r_t)
    lo[:, 1:] = g.idy2 * a[:, :-1] - g.idxy * b[:-1, :] + g.idxy * b[1:, :]
    hi[:, :-1] = g.idy2 * a[:, 1:] + g.idxy * b[:-1, :] - g.idxy * b[1:, :]
    a_next = ((a_rhs + dt_rho * (jl + rh) + dtrk * dd * (lo + hi))
              / (r_t(1.0) + r_t(2.0) * dtrk * g.idy2)).astype(r_t)

    # ---- b edges (all i, j < Ny-1) : td.h:412-457
    rh = np.zeros((Nx, Ny - 1), dtype=r_t)
    dd = np.ones((Nx, Ny - 1), dtype=r_t)
    rh[0, :] = -r_t(2.0) * kappa2 * H * g.idx
    dd[0, :] = 2
    rh[Nx - 1, :] = r_t(2.0) * kappa2 * H * g.idx
    dd[Nx - 1, :] = 2
    on = (mp | pp)[:, :-1]
    jl = np.where(on, g.idy * _js(psi[:, :-1], g.dy * b_ph, psi[:, 1:]), 0).astype(r_t)
    lo = np.zeros((Nx, Ny - 1), dtype=r_t)
    hi = np.zeros((Nx, Ny - 1), dtype=r_t)
    lo[1:, :] = g.idx2 * b[:-1, :] - g.idxy * a[:, :-1] + g.idxy * a[:, 1:]
    hi[:-1, :] = g.idx2 * b[1:, :] + g.idxy * a[:, :-1] - g.idxy * a[:, 1:]
    b_next = ((b_rhs + dt_rho * (jl + rh) + dtrk * dd * (lo + hi))
              / (r_t(1.0) + r_t(2.0) * dtrk * g.idx2)).astype(r_t)

    r = max(np.abs(a_next - a).max(), np.abs(b_next - b).max())
    return a_next, b_next, r_t(r), a_rhs, b_rhs


def td_a_solve(g, dt, kappa, sigma, H, mt, psi, a, b, stop_eps=1e-6, langevin_c=0.0, rand_t=1,
               max_sweeps=1024):
    """A-solve driver (svirl/solvers/td.py:252-325) INCLUDING quirk Q1: the
    phase argument is the Python object holding A at the start of the solve (X);
    buffers ping-pong X<->Y, so on sweep s the phase comes from iterate s-(s%2)."""
    r_t = g.dtype
    kappa2 = r_t(r_t(kappa) ** 2)                      # params.py:101-105
    rho = r_t(1.0 / sigma)                             # params.py:117
    a_rhs, b_rhs = a.copy(), b.copy()
    ca, cb = a, b
    pa, pb = a, b                                      # content of buffer X
    for s in range(max_sweeps):
        if s % 2 == 0:
            pa, pb = ca, cb                            # X is the input
        na, nb, r, a_rhs, b_rhs = a_sweep(g, dt, kappa2, rho, H, mt, psi, pa, pb, a_rhs, b_rhs, ca, cb,
                                          langevin_c, s, rand_t)
        ca, cb = na, nb
        if stop_test(r, stop_eps, r_t):
            break
    return ca, cb, s + 1


def td_run(g, dt, Nt, eps, mt, kappa, sigma, H, psi, a, b, stop_psi=1e-6, stop_A=1e-6,
           langevin_psi=0.0, langevin_A=0.0, rand_t=1, counts=None):
    """TD outer loop (svirl/solvers/td.py:342-367).  kappa=inf => psi only.
    rand_t increments after every psi-solve and every A-solve (td.py:204, 313)."""
    solveA = not np.isposinf(kappa)
    r_t = g.dtype
    stop_psi = max(stop_psi, 1e-6 if r_t is np.float32 else 1e-12)      # td.py:51-66
    stop_A = max(stop_A, 1e-6 if r_t is np.float32 else 1e-12)
    for _ in range(Nt):
        psi, ns = td_psi_solve(g, dt, eps, mt, a, b, psi, stop_psi, langevin_psi, rand_t)
        rand_t = (rand_t + 1) & 0xffffffff
        na = 0
        if solveA:
            a, b, na = td_a_solve(g, dt, kappa, sigma, H, mt, psi, a, b, stop_A, langevin_A, rand_t)
            rand_t = (rand_t + 1) & 0xffffffff
        if counts is not None:
            counts.append((ns, na))
    return psi, a, b, rand_t


# ----------------------------------------------------------------------------------------------
# fixed vortices (svirl/vars/fixed_vortices.py:176-246; svirl/solvers/td.py:120-155, 207-216, 252-325)
# ----------------------------------------------------------------------------------------------

def snap_vortices(g, vx, vy, vv, correction="cell centers"):
    """Integer vorticity, positions on cell centres / vertices (fixed_vortices.py:114-127)."""
    r_t = g.dtype
    vx, vy, vv = [np.asarray(v, dtype=r_t) for v in (vx, vy, vv)]
    vv = np.round(vv)
    if correction == "cell centers":
        vx = g.dx * (np.round(vx / g.dx + 0.5) - 0.5)
        vy = g.dy * (np.round(vy / g.dy + 0.5) - 0.5)
    elif correction == "vertices":
        vx, vy = g.dx * np.round(vx / g.dx), g.dy * np.round(vy / g.dy)
    return vx, vy, vv


def irregular_potential(g, vx, vy, vv):
    """(a_i, b_i) = sum_k v_k * lattice gradient of atan2(y - y_k, x - x_k) (fixed_vortices.py:192-206)."""
    r_t = g.dtype
    xg = np.repeat((g.dx * np.arange(g.Nx, dtype=r_t))[:, None], g.Ny, axis=1)       # mesh/grid.py:35-60
    yg = np.repeat((g.dy * np.arange(g.Ny, dtype=r_t))[None, :], g.Nx, axis=0)
    ai = np.zeros((g.Nx - 1, g.Ny), dtype=r_t)
    bi = np.zeros((g.Nx, g.Ny - 1), dtype=r_t)
    for x0, y0, v in zip(vx, vy, vv):
        th = np.arctan2(yg - y0, xg - x0)
        th -= th[0, 0]
        ai += v * g.idx * np.diff(th, axis=0)
        bi += v * g.idy * np.diff(th, axis=1)
    return ai, bi


def phase_lock_list(g, vx, vy, radius):
    """The reference's lock list: i-indices of the nodes within `radius` of a fixed vortex
    (np.where(mask)[0], fixed_vortices.py:236), later used as FLAT node numbers."""
    r_t = g.dtype
    xg = np.repeat((g.dx * np.arange(g.Nx, dtype=r_t))[:, None], g.Ny, axis=1)
    yg = np.repeat((g.dy * np.arange(g.Ny, dtype=r_t))[None, :], g.Nx, axis=0)
    m = np.zeros((g.Nx, g.Ny), dtype=bool)
    for x0, y0 in zip(vx, vy):
        m |= np.square(xg - x0) + np.square(yg - y0) <= np.square(radius)
    return np.where(m)[0].astype(np.int32)


def _fold_flat(g, xa, xb, ya, yb, sign, n_flat):
    """x[0:n_flat] += sign*y[0:n_flat] on the PACKED edge array (a then b, x fastest): xpy_r / xmy_r are
    launched with N = Nx*Ny on Na+Nb entries (td.py:128,151,260,322), so all of a and the first
    Nx*Ny - Na = Ny entries of b change (quirk Q5)."""
    Na = (g.Nx - 1) * g.Ny
    fa, fb = xa.T.reshape(-1).copy(), xb.T.reshape(-1).copy()
    ga, gb = ya.T.reshape(-1), yb.T.reshape(-1)
    na = min(n_flat, Na)
    fa[:na] += sign * ga[:na]
    nb = max(0, min(n_flat - Na, fb.size))
    fb[:nb] += sign * gb[:nb]
    return fa.reshape(g.Ny, g.Nx - 1).T.astype(g.dtype), fb.reshape(g.Ny - 1, g.Nx).T.astype(g.dtype)


def _phase_lock(g, psi, lock_ns):
    """psi[n] <- |psi[n]| on the flat node list (svirl/cuda/td.h:296-307)."""
    if lock_ns is None or not lock_ns.size:
        return psi
    flat = psi.T.reshape(-1).copy()
    flat[lock_ns] = np.abs(flat[lock_ns])
    return flat.reshape(g.Ny, g.Nx).T.astype(g.ctype)


def td_psi_run_fixed(g, dt, Nt, eps, mt, psi, a, b, ai, bi, lock_ns=None, stop_psi=1e-6, rand_t=1, counts=None):
    """td(eqn='order_parameter') with fixed vortices (svirl/solvers/td.py:221-231): the irregular
    potential is folded in ONCE before the loop and folded out after EVERY step (reference quirk).
    Returns psi, a, b, rand_t."""
    r_t = g.dtype
    stop_psi = max(stop_psi, 1e-6 if r_t is np.float32 else 1e-12)
    N = g.Nx * g.Ny
    a, b = _fold_flat(g, a, b, ai, bi, +1, N)
    for _ in range(Nt):
        psi, ns = td_psi_solve(g, dt, eps, mt, a, b, psi, stop_psi, 0.0, rand_t)
        rand_t = (rand_t + 1) & 0xffffffff
        psi = _phase_lock(g, psi, lock_ns)
        a, b = _fold_flat(g, a, b, ai, bi, -1, N)
        if counts is not None:
            counts.append((ns, 0))
    return psi, a, b, rand_t


def td_run_fixed(g, dt, Nt, eps, mt, kappa, sigma, H, psi, a, b, ai, bi, lock_ns=None, stop_psi=1e-6,
                 stop_A=1e-6, rand_t=1, counts=None):
    """TD outer loop with fixed vortices (td.py:342-367 with :120-155, :207-216, :252-325).  Per step:
    A += A_i (first N entries); psi-solve; phase lock; A -= A_i; A_i += A; A-solve with the link phase
    taken from A_i (a separate buffer: no Q1 aliasing); A_i -= A_new.  Returns psi, a, b, ai, bi, rand_t."""
    solveA = not np.isposinf(kappa)
    r_t = g.dtype
    stop_psi = max(stop_psi, 1e-6 if r_t is np.float32 else 1e-12)
    stop_A = max(stop_A, 1e-6 if r_t is np.float32 else 1e-12)
    N = g.Nx * g.Ny
    for _ in range(Nt):
        a, b = _fold_flat(g, a, b, ai, bi, +1, N)
        psi, ns = td_psi_solve(g, dt, eps, mt, a, b, psi, stop_psi, 0.0, rand_t)
        rand_t = (rand_t + 1) & 0xffffffff
        psi = _phase_lock(g, psi, lock_ns)
        a, b = _fold_flat(g, a, b, ai, bi, -1, N)
        na = 0
        if solveA:
            ai, bi = _fold_flat(g, ai, bi, a, b, +1, N)
            kappa2, rho = r_t(r_t(kappa) ** 2), r_t(1.0 / sigma)
            a_rhs, b_rhs = a.copy(), b.copy()
            ca, cb = a, b
            for s in range(1024):
                ca_, cb_, r, a_rhs, b_rhs = a_sweep(g, dt, kappa2, rho, H, mt, psi, ai, bi, a_rhs, b_rhs, ca, cb, 0.0, s, rand_t)
                ca, cb = ca_, cb_
                if stop_test(r, stop_A, r_t):
                    break
            a, b, na = ca, cb, s + 1
            rand_t = (rand_t + 1) & 0xffffffff
            ai, bi = _fold_flat(g, ai, bi, a, b, -1, N)
        if counts is not None:
            counts.append((ns, na))
    return psi, a, b, ai, bi, rand_t


# ----------------------------------------------------------------------------------------------
# E3: free energy (svirl/cuda/observables.h:251-362 + observables.py:124-149)
# ----------------------------------------------------------------------------------------------

def _du_weights(g, mt):
    """DU weights (svirl/cuda/cg.h:59-64)."""
    r_t = g.dtype
    mm, mp, pm, pp = node_flags(g, mt)
    f = lambda x: x.astype(r_t)
    wW = r_t(0.5) * (f(mm) + f(mp))
    wE = r_t(0.5) * (f(pm) + f(pp))
    wS = r_t(0.5) * (f(mm) + f(pm))
    wN = r_t(0.5) * (f(mp) + f(pp))
    gw = r_t(0.25) * (wW + wE + wS + wN)
    return (mm, mp, pm, pp), (wW, wE, wS, wN), gw


def _sumA(x, y):
    if x is None:
        return y
    if y is None:
        return x
    return x + y


def _g_grad(psi0, ph, psi1):
    """|psi1 U(ph) - psi0|^2 (observables.h:239-247)."""
    c, s = np.cos(ph), np.sin(ph)
    re = psi1.real * c + psi1.imag * s - psi0.real
    im = psi1.imag * c - psi1.real * s - psi0.imag
    return re * re + im * im


def _cellB(g, a, b):
    return g.idx * (b[1:, :] - b[:-1, :]) - g.idy * (a[:, 1:] - a[:, :-1])


def free_energy_density(g, kappa2, eps, H, mt, psi, ae, be, a, b):
    """Per-node pseudo-density before the dx*dy factor; returns (Nx,Ny) array."""
    r_t = g.dtype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    active = mm | mp | pm | pp
    eps_n = np.asarray(eps, dtype=r_t)
    p2 = (psi.real ** 2 + psi.imag ** 2).astype(r_t)
    e = np.where(active, gw * (r_t(0.5) * p2 - eps_n) * p2, 0).astype(r_t)
    A_a, A_b = _sumA(ae, a), _sumA(be, b)
    onE = (pm | pp)[:-1, :]
    e[:-1, :] += np.where(onE, wE[:-1, :] * g.idx2 * _g_grad(psi[:-1, :], g.dx * A_a, psi[1:, :]), 0).astype(r_t)
    onN = (mp | pp)[:, :-1]
    e[:, :-1] += np.where(onN, wN[:, :-1] * g.idy2 * _g_grad(psi[:, :-1], g.dy * A_b, psi[:, 1:]), 0).astype(r_t)
    if kappa2 > 0.0:
        dB = -r_t(H)
        if ae is not None:
            dB = dB + _cellB(g, ae, be)
        if a is not None:
            dB = dB + _cellB(g, a, b)
        e[:-1, :-1] += (r_t(kappa2) * dB * dB).astype(r_t)
    return e


def free_energy(g, kappa2, eps, H, mt, psi, ae, be, a, b):
    """G = dx dy sum(e) (observables.h:358-361).  kappa2 = -1 means kappa = inf."""
    e = free_energy_density(g, kappa2, eps, H, mt, psi, ae, be, a, b)
    return g.dtype(g.dx * g.dy * e.sum(dtype=np.float64))


# ----------------------------------------------------------------------------------------------
# E4 / E5: Jacobians (svirl/cuda/cg.h:5-12, 16-121, 125-301)
# ----------------------------------------------------------------------------------------------

def _grad_jac(psi0, ph, psi1):
    c, s = np.cos(ph), np.sin(ph)
    return 2 * (psi0 - ((psi1.real * c + psi1.imag * s) + 1j * (psi1.imag * c - psi1.real * s)))


def jacobian_psi(g, kappa2, eps, H, mt, psi, ae, be, a, b):
    """dG/dRe psi + i dG/dIm psi, (Nx,Ny) complex (cg.h:58-120)."""
    r_t, c_t = g.dtype, g.ctype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    active = mm | mp | pm | pp
    eps_n = np.asarray(eps, dtype=r_t)
    A_a = _pad_a(g, _sumA(ae, a))
    A_b = _pad_b(g, _sumA(be, b))
    p = psi.real ** 2 + psi.imag ** 2 - eps_n
    jac = 2 * gw * p * psi
    jac = jac + np.where(mm | mp, wW * g.idx2 * _grad_jac(psi, -g.dx * _shift(A_a, -1, 0), _shift(psi, -1, 0)), 0)
    jac = jac + np.where(pm | pp, wE * g.idx2 * _grad_jac(psi, g.dx * A_a, _shift(psi, 1, 0)), 0)
    jac = jac + np.where(mm | pm, wS * g.idy2 * _grad_jac(psi, -g.dy * _shift(A_b, 0, -1), _shift(psi, 0, -1)), 0)
    jac = jac + np.where(mp | pp, wN * g.idy2 * _grad_jac(psi, g.dy * A_b, _shift(psi, 0, 1)), 0)
    jac = np.where(active, jac, 0)
    return (g.dx * g.dy * jac).astype(c_t)


def _curlcurl(g, H, ae, be, a, b):
    """Shared magnetic stencil of jacobian_A (cg.h:176-217, 240-282) and
    current_density (observables.h:66-155), before the kappa2 factor: quirk Q10."""
    r_t = g.dtype
    Nx, Ny = g.Nx, g.Ny
    H = r_t(H)
    # a edges
    ca = np.zeros((Nx - 1, Ny), dtype=r_t)
    dd = np.ones((Nx - 1, Ny), dtype=r_t)
    ca[:, 0] -= r_t(2.0) * H * g.idy
    dd[:, 0] = 2
    ca[:, Ny - 1] += r_t(2.0) * H * g.idy
    dd[:, Ny - 1] = 2

    def nb_a(x, y):
        lo = np.zeros((Nx - 1, Ny), dtype=r_t)
        hi = np.zeros((Nx - 1, Ny), dtype=r_t)
        lo[:, 1:] = -g.idy2 * x[:, :-1] + g.idxy * y[:-1, :] - g.idxy * y[1:, :]
        hi[:, :-1] = -g.idy2 * x[:, 1:] - g.idxy * y[:-1, :] + g.idxy * y[1:, :]
        return lo + hi

    if ae is not None:
        ca = ca + r_t(2.0) / dd * g.idy2 * ae + nb_a(ae, be)
    if a is not None:
        ca = ca + r_t(2.0) * g.idy2 * a + dd * nb_a(a, b)
    # b edges
    cb = np.zeros((Nx, Ny - 1), dtype=r_t)
    dd = np.ones((Nx, Ny - 1), dtype=r_t)
    cb[0, :] += r_t(2.0) * H * g.idx
    dd[0, :] = 2
    cb[Nx - 1, :] -= r_t(2.0) * H * g.idx
    dd[Nx - 1, :] = 2

    def nb_b(x, y):
        lo = np.zeros((Nx, Ny - 1), dtype=r_t)
        hi = np.zeros((Nx, Ny - 1), dtype=r_t)
        lo[1:, :] = -g.idx2 * y[:-1, :] + g.idxy * x[:, :-1] - g.idxy * x[:, 1:]
        hi[:-1, :] = -g.idx2 * y[1:, :] - g.idxy * x[:, :-1] + g.idxy * x[:, 1:]
        return lo + hi

    if ae is not None:
        cb = cb + r_t(2.0) / dd * g.idx2 * be + nb_b(ae, be)
    if a is not None:
        cb = cb + r_t(2.0) * g.idx2 * b + dd * nb_b(a, b)
    return ca.astype(r_t), cb.astype(r_t)


def jacobian_A(g, kappa2, H, mt, psi, ae, be, a, b):
    """(dG/da, dG/db) (cg.h:174-300)."""
    r_t = g.dtype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    ca, cb = _curlcurl(g, H, ae, be, a, b)
    ja = r_t(kappa2) * ca
    jb = r_t(kappa2) * cb
    A_a, A_b = _sumA(ae, a), _sumA(be, b)
    onE = (pm | pp)[:-1, :]
    ja = ja + np.where(onE, -wE[:-1, :] * g.idx * _js(psi[:-1, :], g.dx * A_a, psi[1:, :]), 0)
    onN = (mp | pp)[:, :-1]
    jb = jb + np.where(onN, -wN[:, :-1] * g.idy * _js(psi[:, :-1], g.dy * A_b, psi[:, 1:]), 0)
    f = r_t(2.0) * g.dx * g.dy
    return (f * ja).astype(r_t), (f * jb).astype(r_t)


# ----------------------------------------------------------------------------------------------
# E6 / E7: line-search polynomial coefficients (svirl/cuda/cg.h:305-474, 478-731)
# ----------------------------------------------------------------------------------------------

def _grad_c(psi0, ph, psi1):
    c, s = np.cos(ph), np.sin(ph)
    return ((psi1.real * c + psi1.imag * s) + 1j * (psi1.imag * c - psi1.real * s)) - psi0


def _coef_eps(eps, r_t):
    """Quirk Q11: both coefficient kernels take ``epsilon_spatial`` but never read it
    (no ``if (epsilon_spatial != NULL)`` in cg.h:315-474 / 478-731), so with a spatial
    linear coefficient they use the scalar argument, which the host sets to 0.0
    (svirl/vars/params.py:79-83).  Reproduced, not fixed."""
    if np.ndim(eps) > 0 and np.size(eps) > 1:
        return r_t(0.0)
    return r_t(np.asarray(eps).reshape(-1)[0])


def _local_terms(g, gw, active, eps_n, psi, dpsi):
    r_t = g.dtype
    p2 = psi.real ** 2 + psi.imag ** 2
    d2 = dpsi.real ** 2 + dpsi.imag ** 2
    tw = 2 * (psi.real * dpsi.real + psi.imag * dpsi.imag)
    z = lambda x: np.where(active, x, 0).sum(dtype=np.float64)
    c0 = z(gw * (r_t(0.5) * p2 - eps_n) * p2)
    c1 = z(gw * tw * (p2 - eps_n))
    c2 = z(gw * (-eps_n * d2 + r_t(0.5) * tw * tw + p2 * d2))
    c3 = z(gw * tw * d2)
    c4 = z(gw * r_t(0.5) * d2 * d2)
    return c0, c1, c2, c3, c4


def coef_psi(g, kappa2, eps, H, mt, psi, dpsi, ae, be, a, b):
    """c0..c4 of G(psi + alpha dpsi) (cg.h:400-467).  NOTE: the magnetic term
    uses only the regular (a,b), not the external part (cg.h:449-455)."""
    r_t = g.dtype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    active = mm | mp | pm | pp
    eps_n = _coef_eps(eps, r_t)
    c = list(_local_terms(g, gw, active, eps_n, psi, dpsi))
    A_a, A_b = _sumA(ae, a), _sumA(be, b)
    for (on, w, i2, d, A, s0, s1) in (
            ((pm | pp)[:-1, :], wE[:-1, :], g.idx2, g.dx, A_a, (slice(0, -1), slice(None)), (slice(1, None), slice(None))),
            ((mp | pp)[:, :-1], wN[:, :-1], g.idy2, g.dy, A_b, (slice(None), slice(0, -1)), (slice(None), slice(1, None)))):
        ph = d * A if A is not None else np.zeros(on.shape, dtype=r_t)
        p0, p1, d0, d1 = psi[s0], psi[s1], dpsi[s0], dpsi[s1]
        z = lambda x: np.where(on, x, 0).sum(dtype=np.float64)
        c[2] += z(w * i2 * _g_grad(d0, ph, d1))
        zz = np.conj(_grad_c(p0, ph, p1)) * _grad_c(d0, ph, d1)
        c[1] += z(w * i2 * 2 * zz.real)
        c[0] += z(w * i2 * _g_grad(p0, ph, p1))
    if kappa2 > 0.0:
        dB = _cellB(g, a, b) - r_t(H)
        c[0] += (r_t(kappa2) * dB * dB).sum(dtype=np.float64)
    return (np.array(c) * float(g.dx * g.dy)).astype(r_t)


def coef(g, kappa2, eps, H, mt, psi, dpsi, ae, be, a, b, da, db):
    """17 coefficients as a (5,5) matrix c[i,j] <-> alpha_psi^i alpha_A^j
    (cg.h:528-701 + cg.py:367-372)."""
    r_t = g.dtype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    active = mm | mp | pm | pp
    eps_n = _coef_eps(eps, r_t)
    C = np.zeros((5, 5), dtype=np.float64)
    C[0, 0], C[1, 0], C[2, 0], C[3, 0], C[4, 0] = _local_terms(g, gw, active, eps_n, psi, dpsi)
    A_a, A_b = _sumA(ae, a), _sumA(be, b)
    for (on, w, i2, d, A, dA, s0, s1) in (
            ((pm | pp)[:-1, :], wE[:-1, :], g.idx2, g.dx, A_a, da, (slice(0, -1), slice(None)), (slice(1, None), slice(None))),
            ((mp | pp)[:, :-1], wN[:, :-1], g.idy2, g.dy, A_b, db, (slice(None), slice(0, -1)), (slice(None), slice(1, None)))):
        ph = d * A if A is not None else np.zeros(on.shape, dtype=r_t)
        dph = d * dA
        dph2 = dph * dph
        p0, p1, d0, d1 = psi[s0], psi[s1], dpsi[s0], dpsi[s1]
        z = lambda x: np.where(on, x, 0).sum(dtype=np.float64)
        C[0, 0] += z(w * i2 * _g_grad(p0, ph, p1))
        zz = np.conj(_grad_c(p0, ph, p1)) * _grad_c(d0, ph, d1)
        C[1, 0] += z(w * i2 * 2 * zz.real)
        C[2, 0] += z(w * i2 * _g_grad(d0, ph, d1))
        Um = _U(-ph)
        for row, zz in ((0, p0 * Um * np.conj(p1)),
                        (1, p0 * Um * np.conj(d1) + d0 * Um * np.conj(p1)),
                        (2, d0 * Um * np.conj(d1))):
            C[row, 1] += z(w * i2 * 2 * zz.imag * dph)
            C[row, 2] += z(w * i2 * zz.real * dph2)
            C[row, 3] += z(-w * i2 / r_t(3.0) * zz.imag * dph2 * dph)
            C[row, 4] += z(-w * i2 / r_t(12.0) * zz.real * dph2 * dph2)
    if kappa2 > 0.0:
        BH = -r_t(H)
        if ae is not None:
            BH = BH + _cellB(g, ae, be)
        if a is not None:
            BH = BH + _cellB(g, a, b)
        dB = _cellB(g, da, db)
        k2 = r_t(kappa2)
        C[0, 0] += (k2 * BH * BH).sum(dtype=np.float64)
        C[0, 1] += (k2 * 2 * BH * dB).sum(dtype=np.float64)
        C[0, 2] += (k2 * dB * dB).sum(dtype=np.float64)
    return (C * float(g.dx * g.dy)).astype(r_t)


def alpha_psi_min(c):
    """Smallest non-negative real root of dP/dalpha (svirl/solvers/cg.py:227-235)."""
    am = np.polynomial.polynomial.polyroots([c[1], 2.0 * c[2], 3.0 * c[3], 4.0 * c[4]])
    am = am[np.isclose(am.imag, 0)].real
    am = am[am >= 0]
    return np.min(am)


def alpha_min(c, alpha0=(0.0, 0.0), tol=1e-8):
    """2-variable quartic minimised by SciPy BFGS (svirl/solvers/cg.py:378-419)."""
    import scipy.optimize
    P = np.polynomial.polynomial
    cj0, cj1 = P.polyder(c, axis=0), P.polyder(c, axis=1)
    f = lambda al: P.polyval2d(al[0], al[1], c)
    j = lambda al: np.array([P.polyval2d(al[0], al[1], cj0), P.polyval2d(al[0], al[1], cj1)])
    return scipy.optimize.minimize(f, x0=np.array(alpha0), jac=j, method="BFGS", tol=tol).x


def _beta(gj, gp):
    """PR+ (svirl/cuda/utils.h:13-70, 140-146)."""
    if np.iscomplexobj(gj):
        num = (gj.real * (gj.real - gp.real) + gj.imag * (gj.imag - gp.imag)).sum(dtype=np.float64)
        den = (gp.real ** 2 + gp.imag ** 2).sum(dtype=np.float64)
    else:
        num = (gj * (gj - gp)).sum(dtype=np.float64)
        den = (gp * gp).sum(dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.float64(num) / np.float64(den)
    return max(q, 0.0) if not np.isnan(q) else 0.0     # fmax(nan, 0) = 0 on the device


class CGState(object):
    """State that survives across cg() calls (quirk Q6, svirl/solvers/cg.py:59-71, 452-477)."""

    def __init__(self):
        self.beta_psi = 0.0
        self.beta_A = 0.0
        self.dir_psi = None
        self.dir_a = None
        self.dir_b = None
        self.jp_psi = None
        self.jp_a = None
        self.jp_b = None


def cg_run(g, n_iter, kappa, eps, H, mt, psi, ae, be, a, b, rtol=1e-6, state=None):
    """Modified NLCG (svirl/solvers/cg.py:238-322 for kappa=inf, :452-551 finite).
    Returns (psi, a, b, energies, state)."""
    r_t, c_t = g.dtype, g.ctype
    solveA = not np.isposinf(kappa)
    kappa2 = r_t(r_t(kappa) ** 2) if solveA else r_t(-1.0)
    st = state or CGState()
    if st.dir_psi is None or not solveA:                # kappa=inf path resets dir (cg.py:264)
        st.dir_psi = np.zeros((g.Nx, g.Ny), dtype=c_t)
    if solveA and st.dir_a is None:
        st.dir_a = np.zeros((g.Nx - 1, g.Ny), dtype=r_t)
        st.dir_b = np.zeros((g.Nx, g.Ny - 1), dtype=r_t)
    energies = []
    for i in range(n_iter):
        jpsi = jacobian_psi(g, kappa2, eps, H, mt, psi, ae, be, a, b)
        if solveA:
            ja, jb = jacobian_A(g, kappa2, H, mt, psi, ae, be, a, b)
        if i > 0:
            st.beta_psi = _beta(jpsi, st.jp_psi)
            if solveA:
                st.beta_A = _beta(np.concatenate([ja.ravel(), jb.ravel()]),
                                  np.concatenate([st.jp_a.ravel(), st.jp_b.ravel()]))
        st.dir_psi = (r_t(st.beta_psi) * st.dir_psi - jpsi).astype(c_t)
        if solveA:
            st.dir_a = (r_t(st.beta_A) * st.dir_a - ja).astype(r_t)
            st.dir_b = (r_t(st.beta_A) * st.dir_b - jb).astype(r_t)
            c = coef(g, kappa2, eps, H, mt, psi, st.dir_psi, ae, be, a, b, st.dir_a, st.dir_b)
            al_psi, al_A = alpha_min(c)
            psi = (r_t(al_psi) * st.dir_psi + psi).astype(c_t)
            a = (r_t(al_A) * st.dir_a + a).astype(r_t)
            b = (r_t(al_A) * st.dir_b + b).astype(r_t)
            st.jp_a, st.jp_b = ja, jb
        else:
            c = coef_psi(g, kappa2, eps, H, mt, psi, st.dir_psi, ae, be, a, b)
            al = alpha_psi_min(c)
            psi = (r_t(al) * st.dir_psi + psi).astype(c_t)
        st.jp_psi = jpsi
        energies.append(free_energy(g, kappa2, eps, H, mt, psi, ae, be, a, b))
        if i > 0 and np.abs(energies[i] / energies[i - 1] - 1.0) < rtol:
            break
    return psi, a, b, energies, st


# ----------------------------------------------------------------------------------------------
# observables (svirl/cuda/observables.h:5-235)
# ----------------------------------------------------------------------------------------------

def magnetic_field(g, ae, be, a, b):
    B = np.zeros((g.Nx - 1, g.Ny - 1), dtype=g.dtype)
    if ae is not None:
        B = B + _cellB(g, ae, be)
    if a is not None:
        B = B + _cellB(g, a, b)
    return B.astype(g.dtype)


def current_density(g, kappa2, H, ae, be, a, b):
    ca, cb = _curlcurl(g, H, ae, be, a, b)
    return (g.dtype(kappa2) * ca).astype(g.dtype), (g.dtype(kappa2) * cb).astype(g.dtype)


def supercurrent_density(g, mt, psi, ae, be, a, b):
    r_t = g.dtype
    (mm, mp, pm, pp), (wW, wE, wS, wN), gw = _du_weights(g, mt)
    A_a, A_b = _sumA(ae, a), _sumA(be, b)
    onE = (pm | pp)[:-1, :]
    jx = np.where(onE, wE[:-1, :] * g.idx * _js(psi[:-1, :], g.dx * A_a, psi[1:, :]), 0).astype(r_t)
    onN = (mp | pp)[:, :-1]
    jy = np.where(onN, wN[:, :-1] * g.idy * _js(psi[:, :-1], g.dy * A_b, psi[:, 1:]), 0).astype(r_t)
    return jx, jy


# ----------------------------------------------------------------------------------------------
# vortex detector (svirl/observables/vortex_detector.py:31-156) -- plain scalar restatement
# ----------------------------------------------------------------------------------------------

def winding(g, H, psi, a, b):
    """Vectorised winding number per cell (vortex_detector.py:62-69)."""
    t = np.angle(psi)
    pi = np.pi
    dx, dy = g.dx, g.dy
    v = -(0.5 / pi) * (
        np.mod(t[1:, :-1] - t[:-1, :-1] - dx * a[:, :-1] + pi, 2.0 * pi)
        + np.mod(t[1:, 1:] - t[1:, :-1] - dy * b[1:, :] + pi, 2.0 * pi)
        + np.mod(t[:-1, 1:] - t[1:, 1:] + dx * a[:, 1:] + pi, 2.0 * pi)
        + np.mod(t[:-1, :-1] - t[:-1, 1:] + dy * b[:-1, :] + pi, 2.0 * pi)
        - 4.0 * pi + dx * dy * H)
    return v


def _find_zero(x1, y1, f1, x2, y2, f2):
    return (f2 * x1 - x2 * f1) / (f2 - f1), (f2 * y1 - y2 * f1) / (f2 - f1)


def _zero_line(p):
    (x1, y1, f1), (x2, y2, f2), (x3, y3, f3), (x4, y4, f4) = p
    out = []
    if f2 * f1 < -1e-10:
        out.append(_find_zero(x2, y2, f2, x1, y1, f1))
    if f3 * f2 < -1e-10:
        out.append(_find_zero(x3, y3, f3, x2, y2, f2))
    if f4 * f3 < -1e-10:
        out.append(_find_zero(x4, y4, f4, x3, y3, f3))
    if f1 * f4 < -1e-10:
        out.append(_find_zero(x1, y1, f1, x4, y4, f4))
    return out


def _intersect(l1x1, l1y1, l1x2, l1y2, l2x1, l2y1, l2x2, l2y2):
    D = (l1x1 - l1x2) * (l2y1 - l2y2) - (l1y1 - l1y2) * (l2x1 - l2x2)
    ph = np.arctan2(D, (l1y1 - l1y2) * (l2y1 - l2y2) - (l1x1 - l1x2) * (l2x1 - l2x2))
    ph = np.mod(np.abs(ph), 0.5 * np.pi)
    if np.abs(ph) > 1e-10:
        ix = ((l1x1 * l1y2 - l1y1 * l1x2) * (l2x1 - l2x2) - (l1x1 - l1x2) * (l2x1 * l2y2 - l2y1 * l2x2)) / D
        iy = ((l1x1 * l1y2 - l1y1 * l1x2) * (l2y1 - l2y2) - (l1y1 - l1y2) * (l2x1 * l2y2 - l2y1 * l2x2)) / D
        return ix, iy, ph
    return np.nan, np.nan, ph


def vortices(g, H, psi, a, b, ai=None, bi=None):
    """Returns (x, y, vorticity) arrays in g.dtype, ordered by ascending cell
    index n = i + (Nx-1) j, exactly like the reference's loop."""
    Nxc = g.Nx - 1
    dx, dy = g.dx, g.dy
    v = winding(g, H, psi, a, b)
    a_ai = a + ai if ai is not None else a
    b_bi = b + bi if bi is not None else b
    out = []
    cand = np.argwhere((np.abs(v) > 0.5) & (np.abs(v - np.round(v)) < 0.1))
    cand = sorted((int(i) + Nxc * int(j), int(i), int(j)) for i, j in cand)
    for n, i, j in cand:
        # the reference's i, j are np.int32 (n % cfg.Nxc with cfg.Nxc an np.int32,
        # vortex_detector.py:23-24), so dx*i promotes to float64 even in fp32 mode
        i, j = np.int32(i), np.int32(j)
        ip, jp = i + 1, j + 1
        x, y = dx * i, dy * j
        ia00, ia0p = dx * a_ai[i, j], dx * a_ai[i, jp]
        ib00, ibp0 = dy * b_bi[i, j], dy * b_bi[ip, j]
        t00 = psi[i, j]
        tp0 = psi[ip, j] * np.exp(-1j * (0.75 * ia00 + 0.25 * (ib00 + ia0p - ibp0)))
        tpp = psi[ip, jp] * np.exp(-1j * (0.5 * (ia00 + ibp0) + 0.5 * (ib00 + ia0p)))
        t0p = psi[i, jp] * np.exp(-1j * (0.75 * ib00 + 0.25 * (ia00 + ibp0 - ia0p)))
        corners = lambda f: ((x, y, f(t00)), (x + dx, y, f(tp0)), (x + dx, y + dy, f(tpp)), (x, y + dy, f(t0p)))
        re_xy, im_xy = _zero_line(corners(np.real)), _zero_line(corners(np.imag))
        if len(re_xy) == 2 and len(im_xy) == 2:
            ix, iy, ph = _intersect(re_xy[0][0], re_xy[0][1], re_xy[1][0], re_xy[1][1],
                                    im_xy[0][0], im_xy[0][1], im_xy[1][0], im_xy[1][1])
            if x - dx < ix < x + 2.0 * dx and y - dy < iy < y + 2.0 * dy:
                out.append([ix, iy, np.round(v[i, j])])
    arr = np.array(out, dtype=g.dtype) if out else np.zeros((0, 3), dtype=g.dtype)
    return arr[:, 0], arr[:, 1], arr[:, 2]


# ----------------------------------------------------------------------------------------------
# initial state (svirl/vars/vars.py:93-109, svirl/vars/params.py:135-170, svirl/mesh/grid.py:51-76)
# ----------------------------------------------------------------------------------------------

def initial_psi(g, level=1.0, seed=None):
    """psi0 = (1 - l u1) exp(i pi l (2 u2 - 1)); u1 = first N legacy-MT draws,
    u2 = next N; flat index n = i + Nx j (x fastest)."""
    N = g.Nx * g.Ny
    if seed is not None:
        np.random.seed(seed)
    data = (1.0 - level * np.random.rand(N)) * np.exp(level * 1.0j * np.pi * (2.0 * np.random.rand(N) - 1.0))
    return np.reshape(data.astype(g.ctype), (g.Ny, g.Nx)).T.copy()


def initial_A(g, H):
    """Symmetric gauge on edge midpoints: a = -(y - Ly/2) H / 2, b = +(x - Lx/2) H / 2."""
    r_t = g.dtype
    ya = np.linspace(0.0, g.Ly, num=g.Ny, endpoint=True, dtype=r_t)
    xb = np.linspace(0.0, g.Lx, num=g.Nx, endpoint=True, dtype=r_t)
    _, yg = np.meshgrid(np.zeros(g.Nx - 1, dtype=r_t), ya, indexing="ij")
    xg, _ = np.meshgrid(xb, np.zeros(g.Ny - 1, dtype=r_t), indexing="ij")
    a = np.zeros((g.Nx - 1, g.Ny), dtype=r_t)
    b = np.zeros((g.Nx, g.Ny - 1), dtype=r_t)
    a -= 0.5 * (yg - 0.5 * g.Ly) * r_t(H)
    b += (1.0 - 0.5) * (xg - 0.5 * g.Lx) * r_t(H)
    return a, b
