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
