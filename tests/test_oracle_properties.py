"""CPU: the properties the reference's own acceptance tests assert (svirl tests/at_cg_jacobians.py,
at_cg_coef.py, at_cg_coef_psi.py), checked on the NumPy oracle -- together with the golden fixtures this
pins the oracle to the reference."""
import numpy as np
import pytest

import glnumpy as O

P = np.polynomial.polynomial


def random_case(seed, tiled):
    rs = np.random.RandomState(seed)
    Nx, Ny = 8 + rs.randint(4), 8 + rs.randint(4)
    g = O.Grid(Nx, Ny, 0.5 - 0.1 * rs.rand(), 0.5 - 0.1 * rs.rand(), np.float64)
    kappa2 = np.float64((1.0 + 3.0 * rs.rand()) ** 2)
    H, He = 0.01 + 0.1 * rs.rand(), 0.01 + 0.1 * rs.rand()
    psi = O.initial_psi(g, 0.5, int(rs.randint(1 << 30)))
    mt = rs.rand(Nx - 1, Ny - 1) > 0.3 if tiled else None
    if mt is not None:
        mm, mp, pm, pp = O.node_flags(g, mt)
        psi = psi * (mm | mp | pm | pp)
    a, b = O.initial_A(g, H)              # curl a = H, like the reference's test
    ae, be = O.initial_A(g, He)
    return g, kappa2, H, mt, psi, ae, be, a, b, rs


@pytest.mark.parametrize("seed,tiled", [(1, False), (2, True), (3, True)])
def test_jacobians_equal_finite_differences(seed, tiled):
    g, k2, H, mt, psi, ae, be, a, b, _ = random_case(seed, tiled)
    E = lambda p, x, y: O.free_energy(g, k2, 1.0, H, mt, p, ae, be, x, y)
    E0, h = E(psi, a, b), 3e-9
    active = np.ones(psi.shape, bool)
    if mt is not None:
        mm, mp, pm, pp = O.node_flags(g, mt)
        active = mm | mp | pm | pp
    fd = np.zeros_like(psi)
    for i in range(g.Nx):
        for j in range(g.Ny):
            if not active[i, j]:
                continue                  # psi is kept at zero outside the material
            for unit in (1.0, 1.0j):
                p = psi.copy()
                p[i, j] += unit * h
                fd[i, j] += unit * (E(p, a, b) - E0) / h
    assert np.allclose(O.jacobian_psi(g, k2, 1.0, H, mt, psi, ae, be, a, b), fd, atol=1e-5, rtol=1e-3)
    ja, jb = O.jacobian_A(g, k2, H, mt, psi, ae, be, a, b)
    fa, fb = np.zeros_like(a), np.zeros_like(b)
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            x = a.copy()
            x[i, j] += h
            fa[i, j] = (E(psi, x, b) - E0) / h
    for i in range(b.shape[0]):
        for j in range(b.shape[1]):
            y = b.copy()
            y[i, j] += h
            fb[i, j] = (E(psi, a, y) - E0) / h
    assert np.allclose(ja, fa, atol=1e-5, rtol=1e-3) and np.allclose(jb, fb, atol=1e-5, rtol=1e-3)


@pytest.mark.parametrize("seed,tiled", [(4, False), (5, True)])
def test_line_search_polynomials_reproduce_energy(seed, tiled):
    g, k2, H, mt, psi0, ae, be, a0, b0, rs = random_case(seed, tiled)
    E0 = O.free_energy(g, k2, 1.0, H, mt, psi0, ae, be, a0, b0)
    for r in (0.001, 0.01, 0.1, 0.3):
        for a_psi, b_psi, al in ((1.0, 0.0, 0.0), (0.5, 0.5, 1.0), (0.6976, 0.72, 0.42)):
            c = O.coef(g, k2, 1.0, H, mt, psi0 * a_psi, psi0 * b_psi, ae, be, a0 * (1 - r), b0 * (1 - r), a0 * r, b0 * r)
            tr = lambda o: P.polyval2d(al, 1.0, c * (np.arange(5)[None, :] <= o))
            assert np.isclose(tr(4), E0)
            err = np.abs(np.array([tr(0), tr(2), tr(4)]) - E0)
            err[err < 1e-9] = 0.0
            assert np.all(np.diff(err) < 1e-14)
    # infinite kappa: the quartic in alpha_psi is exact
    kinf = np.float64(-1.0)
    Ei = O.free_energy(g, kinf, 1.0, H, mt, psi0, ae, be, a0, b0)
    c5 = O.coef_psi(g, kinf, 1.0, H, mt, psi0 * 0.6976, psi0 * 0.72, ae, be, a0, b0)
    assert np.isclose(P.polyval(0.42, c5), Ei, rtol=1e-12)
