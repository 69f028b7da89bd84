"""CPU: the host line searches of the CG solver (svirl_b200/solvers/cg.py) without a device: the psi-only
root selection against the oracle's restatement, and the rescue of the 2-D BFGS search on coefficients
measured at 16384^2, where the reference's call runs away."""
import numpy as np

import glnumpy as O
import svirl_b200.config as cfg
from svirl_b200.solvers.cg import CG

C16 = np.array([[-5.569412e-02, -5.033755e-03, 7.153037e-02, 2.296557e-04, -4.050537e-04],
                [-2.623664e-02, -1.681635e-02, 2.434797e-02, 1.979796e-04, -2.142220e-04],
                [2.284921e-02, -3.724204e-03, 5.584273e-03, 3.571101e-05, -4.655317e-05],
                [1.439366e-02, 0, 0, 0, 0], [2.060112e-03, 0, 0, 0, 0]])


def bare_cg(c):
    cg = object.__new__(CG)                 # no device: only the line-search methods are used
    cg._CG__c = np.array(c, dtype=np.float64)
    cg.line_search_rescues = 0
    return cg


def test_guarded_search_rescues_runaway_and_keeps_good_results(monkeypatch):
    monkeypatch.setattr(cfg, "cg_line_search", "reference", raising=False)
    cg = bare_cg(C16 * 16384.0 ** 2)
    with np.errstate(all="ignore"):
        raw = cg._cg_alpha_min()
    assert not (np.all(np.isfinite(raw)) and np.max(np.abs(raw)) < 1e6)
    a = cg._cg_alpha_min_guarded()
    assert cg.line_search_rescues == 1 and np.allclose(a, [0.42212742, 0.07723854], rtol=1e-4)
    ok = bare_cg(C16)
    assert np.array_equal(ok._cg_alpha_min_guarded(), ok._cg_alpha_min()) and ok.line_search_rescues == 0
    assert np.allclose(ok._cg_alpha_min(), O.alpha_min(C16), rtol=0, atol=0)
    monkeypatch.setattr(cfg, "cg_line_search", "normalized", raising=False)
    assert np.allclose(bare_cg(C16 * 1e9)._cg_alpha_min_guarded(), [0.42212742, 0.07723854], rtol=1e-4)


def test_psi_root_selection_equals_oracle():
    rs = np.random.RandomState(2)
    for _ in range(50):
        c = rs.randn(5)
        c[4] = abs(c[4]) + 0.1              # quartic bounded below, like the energy along a direction
        c[1] = -abs(c[1])                   # descent direction
        cg = bare_cg(c)
        assert cg._cg_alpha_psi_min() == O.alpha_psi_min(c)


def _sample_coefficients():
    """17-coefficient sets of a real CG run (oracle, 48^2, kappa 2), at several scales incl. run-away ones."""
    g = O.Grid(48, 48, 0.5, 0.5, np.float64)
    psi = O.initial_psi(g, 1.0, 1234)
    a, b = O.initial_A(g, 0.1)
    psi, a, b, _ = O.td_run(g, 0.1, 10, 1.0, None, 2.0, 10.0, 0.1, psi, a, b, rand_t=1234)
    cs, orig = [], O.alpha_min

    def wrap(c):
        cs.append(np.array(c))
        return orig(c)
    O.alpha_min = wrap
    try:
        O.cg_run(g, 6, 2.0, 1.0, 0.1, None, psi, np.zeros_like(a), np.zeros_like(b), a, b, rtol=-1.0)
    finally:
        O.alpha_min = orig
    return cs


def test_fast_callables_give_scipy_the_same_bits():
    """_cg_alpha_min evaluates the polynomial with _horner2d instead of numpy's polyval2d: SciPy's BFGS must see
    identical values, i.e. return the identical minimiser after identical iteration / evaluation counts -- also on
    coefficient sets where the search runs away."""
    import scipy.optimize
    from svirl_b200.solvers.cg import _horner2d
    P = np.polynomial.polynomial
    rs = np.random.RandomState(0)
    sets = [c * s for c in _sample_coefficients() for s in (1.0, 7.0e3, 2.9e4)] + [C16, C16 * 16384.0 ** 2]
    for c in sets:
        for _ in range(50):
            x, y = (rs.randn(2) * rs.choice([1e-3, 1.0, 1e3])).tolist()
            assert P.polyval2d(x, y, c) == _horner2d(x, y, c.tolist())
        cg = bare_cg(c)
        with np.errstate(all="ignore"):
            got, want = cg._cg_alpha_min(), cg._cg_alpha_min_numpy()
        assert np.array_equal(got, want, equal_nan=True)
    # iteration and evaluation counts (what SciPy did, not only where it ended)
    c = sets[1]
    cj0, cj1 = P.polyder(c, axis=0), P.polyder(c, axis=1)
    r1 = scipy.optimize.minimize(lambda al: P.polyval2d(al[0], al[1], c), x0=np.zeros(2), method="BFGS", tol=1e-8,
                                 jac=lambda al: np.array([P.polyval2d(al[0], al[1], cj0), P.polyval2d(al[0], al[1], cj1)]))
    C, C0, C1 = c.tolist(), cj0.tolist(), cj1.tolist()
    r2 = scipy.optimize.minimize(lambda al: np.float64(_horner2d(float(al[0]), float(al[1]), C)), x0=np.zeros(2),
                                 method="BFGS", tol=1e-8,
                                 jac=lambda al: np.array([_horner2d(float(al[0]), float(al[1]), C0),
                                                          _horner2d(float(al[0]), float(al[1]), C1)]))
    assert (r1.nit, r1.nfev, r1.njev) == (r2.nit, r2.nfev, r2.njev) and np.array_equal(r1.x, r2.x)


def test_native_line_search_finds_the_reference_minimum(monkeypatch):
    """cfg.cg_line_search = 'native' (svl_cg_line_search, damped Newton on c / max|c|): the same minimiser as the
    reference's BFGS where that converges (to its 1e-8 termination error), a finite descent step where it runs away."""
    monkeypatch.setattr(cfg, "cg_line_search", "native", raising=False)
    P = np.polynomial.polynomial
    for c in _sample_coefficients():
        want = bare_cg(c)._cg_alpha_min_numpy()
        got = bare_cg(c)._cg_alpha_min_guarded()
        assert np.allclose(got, want, rtol=2e-6, atol=1e-9), (got, want)
        assert P.polyval2d(got[0], got[1], c) <= P.polyval2d(want[0], want[1], c) + 1e-12 * abs(c[0, 0])
        assert np.allclose(bare_cg(c * 3.0e8)._cg_alpha_min_guarded(), got, rtol=1e-12)      # scale invariant
    a = bare_cg(C16 * 16384.0 ** 2)._cg_alpha_min_guarded()
    assert np.allclose(a, [0.42212742, 0.07723854], rtol=1e-4)
