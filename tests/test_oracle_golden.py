"""CPU: the NumPy oracle against the reference's own outputs (tests/golden, made by
oracle/make_golden.py from the unmodified reference running behind the SIMT shim)."""
import numpy as np
import pytest

import glnumpy as O
from conftest import load_golden, golden_inputs

TD_CASES = ["td_f64_k5", "td_f64_k2_tiled_eps", "td_f64_kinf", "td_f64_k3_langevin",
            "td_f32_kinf_tiled", "td_f32_k2_tiled_eps", "td_f32_k3_langevin"]


def _grid(meta_or_shape, dtype, dx=0.5, dy=0.4):
    Nx, Ny = meta_or_shape
    return O.Grid(Nx, Ny, dx, dy, dtype)


@pytest.mark.parametrize("name", TD_CASES)
def test_td_trajectory(name):
    d = load_golden(name)
    m = d["meta"]
    dtype = np.dtype(m["dtype"]).type
    g = O.Grid(m["Nx"], m["Ny"], m["dx"], m["dy"], dtype)
    mt, eps = golden_inputs(d)
    # initial state is reproducible from the seed alone
    assert np.array_equal(O.initial_psi(g, 1.0, m["random_seed"]), d["psi0"])
    a0, b0 = O.initial_A(g, m["homogeneous_external_field"])
    assert np.array_equal(a0, d["a0"]) and np.array_equal(b0, d["b0"])
    counts = []
    psi, a, b, rt = O.td_run(g, 0.1, m["Nt"], eps, mt, m.get("gl_parameter", np.inf),
                             m.get("normal_conductivity", 1.0), m["homogeneous_external_field"],
                             d["psi0"], d["a0"], d["b0"],
                             langevin_psi=m.get("order_parameter_Langevin_coefficient", 0.0),
                             langevin_A=m.get("vector_potential_Langevin_coefficient", 0.0),
                             rand_t=m["random_seed"], counts=counts)
    ns, na = sum(c[0] for c in counts), sum(c[1] for c in counts)
    assert rt == int(d["rand_t"])
    if dtype is np.float64:
        assert (ns, na) == (int(d["sweeps_psi"]), int(d["sweeps_A"]))
        tol = 1e-12
    else:
        assert abs(ns - int(d["sweeps_psi"])) <= 0.05 * ns and abs(na - int(d["sweeps_A"])) <= 0.05 * na + 1
        tol = 1e-4
    assert np.abs(psi - d["psi1"]).max() < tol
    assert np.abs(a - d["a1"]).max() < tol and np.abs(b - d["b1"]).max() < tol
    # observables on the reference's end state
    k2 = dtype(dtype(m["gl_parameter"]) ** 2) if "gl_parameter" in m else dtype(-1.0)
    H = m["homogeneous_external_field"]
    E = O.free_energy(g, k2, eps, H, mt, d["psi1"], d["ae"], d["be"], d["a1"], d["b1"])
    assert abs(E - d["obs_E"]) <= (1e-13 if dtype is np.float64 else 2e-6) * abs(d["obs_E"])
    rt_ = 1e-13 if dtype is np.float64 else 1e-5
    assert np.allclose(O.magnetic_field(g, d["ae"], d["be"], d["a1"], d["b1"]), d["obs_B"], rtol=rt_, atol=rt_)
    jx, jy = O.supercurrent_density(g, mt, d["psi1"], d["ae"], d["be"], d["a1"], d["b1"])
    assert np.allclose(jx, d["obs_jsx"], rtol=rt_, atol=rt_) and np.allclose(jy, d["obs_jsy"], rtol=rt_, atol=rt_)
    if "obs_jx" in d:
        jx, jy = O.current_density(g, k2, H, d["ae"], d["be"], d["a1"], d["b1"])
        assert np.allclose(jx, d["obs_jx"], rtol=rt_, atol=1e-12 if dtype is np.float64 else 1e-4)
        assert np.allclose(jy, d["obs_jy"], rtol=rt_, atol=1e-12 if dtype is np.float64 else 1e-4)
    vx, vy, vv = O.vortices(g, H, d["psi1"], d["a1"], d["b1"])
    assert np.array_equal(vx, d["obs_vx"]) and np.array_equal(vy, d["obs_vy"]) and np.array_equal(vv, d["obs_vv"])


@pytest.mark.parametrize("name", ["kernels_f64_k3_ext", "kernels_f64_kinf", "kernels_f32_k3_ext", "kernels_f32_kinf"])
def test_kernels(name):
    d = load_golden(name)
    dtype = d["a"].dtype.type
    Nx, Ny = d["psi"].shape
    g = O.Grid(Nx, Ny, 0.5, 0.4, dtype)
    mt, eps = golden_inputs(d)
    k2, H = d["kappa2"], d["H"]
    tol = 1e-13 if dtype is np.float64 else 2e-5
    args = (mt, d["psi"], d["ae"], d["be"], d["a"], d["b"])
    E = O.free_energy(g, k2, eps, H, *args)
    assert abs(E - d["E"]) <= tol * abs(d["E"])
    jp = O.jacobian_psi(g, k2, eps, H, *args)
    assert np.abs(jp - d["jac_psi"]).max() <= tol * np.abs(d["jac_psi"]).max()
    c5 = O.coef_psi(g, k2, eps, H, mt, d["psi"], d["dpsi"], d["ae"], d["be"], d["a"], d["b"])
    assert np.abs(c5 - d["coef5"]).max() <= tol * np.abs(d["coef5"]).max()
    if "coef17" in d:
        ja, jb = O.jacobian_A(g, k2, H, *args)
        s = max(np.abs(d["jac_a"]).max(), np.abs(d["jac_b"]).max())
        assert np.abs(ja - d["jac_a"]).max() <= 10 * tol * s and np.abs(jb - d["jac_b"]).max() <= 10 * tol * s
        c = O.coef(g, k2, eps, H, mt, d["psi"], d["dpsi"], d["ae"], d["be"], d["a"], d["b"], d["da"], d["db"])
        assert np.abs(c - d["coef17"]).max() <= tol * np.abs(d["coef17"]).max()


@pytest.mark.parametrize("name", ["cg_f64_kinf", "cg_f32_kinf_tiled"])
def test_cg_psi_trajectory(name):
    """kappa = inf: polyroots line search is well conditioned -> whole trajectory matches."""
    d = load_golden(name)
    dtype = d["a0"].dtype.type
    Nx, Ny = d["psi0"].shape
    g = O.Grid(Nx, Ny, 0.5, 0.4, dtype)
    mt, eps = golden_inputs(d)
    psi, a, b, E1, st = O.cg_run(g, 25, np.inf, eps, d["H"], mt, d["psi0"], d["ae"], d["be"], d["a0"], d["b0"])
    if dtype is np.float64:
        assert len(E1) == len(d["E1"])
        assert np.allclose(E1, d["E1"], rtol=1e-12)
        assert np.abs(psi - d["psi1"]).max() < 1e-11
        psi, a, b, E2, st = O.cg_run(g, 5, np.inf, eps, d["H"], mt, psi, d["ae"], d["be"], a, b, state=st)
        assert np.allclose(E2, d["E2"], rtol=1e-12) and np.abs(psi - d["psi2"]).max() < 1e-11
    else:
        n = min(len(E1), len(d["E1"]))
        assert np.allclose(E1[:n], d["E1"][:n], rtol=2e-3)


@pytest.mark.parametrize("name", ["cg_f64_k2", "cg_f64_k2_tiled_eps"])
def test_cg_full_first_iterations(name):
    """Finite kappa: SciPy BFGS (tol 1e-8) amplifies 1e-15 coefficient noise to ~1e-9 in
    alpha after a few iterations, so only the first iterations are compared tightly."""
    d = load_golden(name)
    Nx, Ny = d["psi0"].shape
    g = O.Grid(Nx, Ny, 0.5, 0.4, np.float64)
    mt, eps = golden_inputs(d)
    psi, a, b, E, st = O.cg_run(g, 3, float(d["kappa"]), eps, d["H"], mt, d["psi0"], d["ae"], d["be"], d["a0"], d["b0"])
    assert np.allclose(E, d["E1"][:3], rtol=1e-9)


def test_cfg1_readme_200_steps():
    d = load_golden("cfg1_td200")
    m = d["meta"]
    g = O.Grid(129, 129, 0.5, 0.5, np.float64)
    assert np.array_equal(O.initial_psi(g, 1.0, 1234), d["psi0"])
    counts = []
    psi, a, b, _ = O.td_run(g, 0.1, 60, 1.0, None, 5.0, 200.0, 0.1, d["psi0"], d["a0"], d["b0"], rand_t=1234, counts=counts)
    # 60 of the 200 steps on the CPU (time budget); the reference's first 60 steps are not stored, so
    # check self-consistency of sweep counts' scale and run the stored-state observables instead
    assert 1000 < sum(c[0] for c in counts) < 2500
    E = O.free_energy(g, np.float64(25.0), 1.0, 0.1, None, d["psi1"], d["ae"], d["be"], d["a1"], d["b1"])
    assert abs(E - d["obs_E"]) < 1e-12 * abs(E)
    vx, vy, vv = O.vortices(g, 0.1, d["psi1"], d["a1"], d["b1"])
    assert vx.size == 58 and np.array_equal(vx, d["obs_vx"]) and np.array_equal(vy, d["obs_vy"])


def test_stop_rule_is_r_lt_eps():
    for dt_ in (np.float32, np.float64):
        eps = dt_(1e-6)
        assert O.stop_test(dt_(0.99e-6), eps, dt_)
        assert not O.stop_test(dt_(1.01e-6), eps, dt_)
        assert not O.stop_test(dt_(1.0), eps, dt_)


def test_hash_rng_known_values():
    # Thomas Wang hash: fixed integer arithmetic, checked against a scalar transcription
    def wang(s):
        s &= 0xffffffff
        s = ((s ^ 61) ^ (s >> 16)) & 0xffffffff
        s = (s * 9) & 0xffffffff
        s = s ^ (s >> 4)
        s = (s * 0x27d4eb2d) & 0xffffffff
        return s ^ (s >> 15)
    n = np.arange(0, 5000, 7, dtype=np.uint32)
    assert np.array_equal(O.rand_hash(n), np.array([wang(int(x)) for x in n], dtype=np.uint32))
