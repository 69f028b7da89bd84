"""GPU: the CUDA path (through GLSolver and the C ABI) against the reference's outputs stored in
tests/golden (made from the unmodified reference, see oracle/make_golden.py) and against the
NumPy oracle on the same seeded inputs.  Tolerances: fp64 1e-10 relative, fp32 1e-4."""
import numpy as np
import pytest

from conftest import load_golden, golden_inputs, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

TD_CASES = ["td_f64_k5", "td_f64_k2_tiled_eps", "td_f64_kinf", "td_f64_k3_langevin",
            "td_f32_kinf_tiled", "td_f32_k2_tiled_eps", "td_f32_k3_langevin"]


def make_solver(m, d, **extra):
    from svirl_b200 import GLSolver
    kw = {k: v for k, v in m.items() if k not in ("Nt", "dtype")}
    kw["dtype"] = np.dtype(m["dtype"]).type
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    if "eps" in d and d["eps"].size > 1:
        kw["linear_coefficient"] = d["eps"]
    kw.update(extra)
    return GLSolver(**kw)


def relerr(x, ref):
    return np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300)


# psi-sweep kernel variants: plain per-node kernel, streaming kernel (TMA staging) at several
# temporal-blocking depths, and the streaming kernel with plain-load staging
KERNELS = [("plain", 0, 1, 1), ("stream_k4_tma", 1, 4, 1), ("stream_k1_tma", 1, 1, 1), ("stream_k3_tma", 1, 3, 1),
           ("stream_k6_tma", 1, 6, 1), ("stream_k4_ldg", 1, 4, 0),
           ("tile_k4", 2, 4, 1), ("tile_k1", 2, 1, 1), ("tile_k3", 2, 3, 1), ("tile_k8", 2, 8, 1),
           ("tile_k4_plainA", 2, 4, 1, 0),        # 5th entry: A-sweep kernel (0 per-node, default 1 = pair tile kernel)
           ("small", -1, 0, 0)]                   # the single-launch cluster kernel for small grids (td_small.cu)


def set_kernel(gl, kernel):
    _, pk, k, tma = kernel[:4]
    gl.par.set_option("graphs", 1 if pk < 0 else 0)      # the fixtures are small grids: choose the path explicitly
    if pk < 0:
        return
    gl.par.set_option("a_kernel", kernel[4] if len(kernel) > 4 else 1)
    gl.par.set_option("psi_kernel", pk)
    gl.par.set_option("psi_k", k)
    gl.par.set_option("tma", tma)


@pytest.mark.parametrize("kernel", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("name", TD_CASES)
def test_td_trajectory(name, kernel):
    d = load_golden(name)
    m = d["meta"]
    f64 = m["dtype"] == "float64"
    gl = make_solver(m, d)
    set_kernel(gl, kernel)
    assert np.array_equal(gl.vars.order_parameter, d["psi0"])
    a0, b0 = gl.vars.vector_potential
    assert np.array_equal(a0, d["a0"]) and np.array_equal(b0, d["b0"])
    gl.solve.td(dt=0.1, Nt=m["Nt"])
    psi = gl.vars.order_parameter
    a, b = gl.vars.vector_potential
    td = gl.solve._td
    tol = 1e-10 if f64 else 1e-4
    if f64:
        assert (td.sweeps_order_parameter, td.sweeps_vector_potential) == (int(d["sweeps_psi"]), int(d["sweeps_A"]))
    else:
        assert abs(td.sweeps_order_parameter - int(d["sweeps_psi"])) <= 0.05 * int(d["sweeps_psi"])
    assert int(td._random_t) == int(d["rand_t"])
    assert relerr(psi, d["psi1"]) < tol
    assert relerr(a, d["a1"]) < tol and relerr(b, d["b1"]) < tol
    # observables of the end state
    E = gl.observables.free_energy
    assert abs(E - d["obs_E"]) < (1e-10 if f64 else 1e-4) * abs(d["obs_E"])
    assert relerr(gl.observables.magnetic_field, d["obs_B"]) < (1e-9 if f64 else 1e-3)
    jsx, jsy = gl.observables.supercurrent_density
    atol = (1e-10 if f64 else 1e-4) * max(np.abs(d["obs_jsx"]).max(), 1e-3)
    assert np.abs(jsx - d["obs_jsx"]).max() < atol and np.abs(jsy - d["obs_jsy"]).max() < atol
    if "obs_jx" in d:
        jx, jy = gl.observables.current_density
        s = max(np.abs(d["obs_jx"]).max(), 1e-3)
        assert np.abs(jx - d["obs_jx"]).max() < (1e-9 if f64 else 2e-3) * s
        assert np.abs(jy - d["obs_jy"]).max() < (1e-9 if f64 else 2e-3) * s
    vx, vy, vv = gl.vortex_detector.vortices
    assert vx.size == d["obs_vx"].size and np.array_equal(vv, d["obs_vv"])
    if f64:
        assert np.allclose(vx, d["obs_vx"], rtol=0, atol=1e-8) and np.allclose(vy, d["obs_vy"], rtol=0, atol=1e-8)


def _set_state(gl, d, psi_key="psi", a_key="a", b_key="b"):
    gl.vars.order_parameter = d[psi_key]
    gl.vars.vector_potential = (d[a_key], d[b_key])
    gl.params.external_vector_potential = (d["ae"], d["be"])


@pytest.mark.parametrize("name", ["kernels_f64_k3_ext", "kernels_f64_kinf", "kernels_f32_k3_ext", "kernels_f32_kinf"])
def test_kernels(name):
    from svirl_b200 import GLSolver
    from svirl_b200.storage import GArray
    d = load_golden(name)
    dtype = d["a"].dtype.type
    f64 = dtype is np.float64
    Nx, Ny = d["psi"].shape
    finite = "coef17" in d
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, homogeneous_external_field=float(d["H"]),
              gl_parameter=float(np.sqrt(d["kappa2"])) if finite else np.inf)
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    if bool(d["eps_is_field"]):
        kw["linear_coefficient"] = d["eps"]
    gl = GLSolver(**kw)
    # kappa2 must be bit-identical to the fixture's
    if finite:
        assert gl.params.gl_parameter_squared_h() == pytest.approx(float(d["kappa2"]), rel=1e-6 if not f64 else 1e-14)
    _set_state(gl, d)
    tol = 1e-12 if f64 else 2e-5
    E = gl.observables.free_energy
    assert abs(E - d["E"]) < tol * abs(d["E"])
    gl.solve._init_cg()
    cg = gl.solve._cg
    jp = gl.unflatten_array(cg._free_energy_jacobian_psi.get())
    assert relerr(jp, d["jac_psi"]) < tol
    dpsi = GArray(like=d["dpsi"])
    if finite:
        jA = cg._free_energy_jacobian_A.get()
        ja, jb = gl.unflatten_a_array(jA[:gl.cfg.Na]), gl.unflatten_b_array(jA[gl.cfg.Na:])
        s = max(np.abs(d["jac_a"]).max(), np.abs(d["jac_b"]).max())
        assert np.abs(ja - d["jac_a"]).max() < 10 * tol * s and np.abs(jb - d["jac_b"]).max() < 10 * tol * s
        dab = GArray(shape=[d["da"].shape, d["db"].shape], dtype=dtype)
        dab.set_vec_h(d["da"], d["db"])
        dab.sync()
        c = np.array(cg._free_energy_conjgrad_coef(dpsi.get_d_obj(), dab.get_d_obj()))
        assert relerr(c, d["coef17"]) < tol
    c5 = np.array(cg._free_energy_conjgrad_coef_psi(dpsi.get_d_obj()))
    c5 = c5[0, :] if c5.ndim == 2 else c5
    assert relerr(c5, d["coef5"]) < tol


@pytest.mark.parametrize("name", ["cg_f64_kinf", "cg_f32_kinf_tiled"])
def test_cg_psi_trajectory(name):
    from svirl_b200 import GLSolver
    d = load_golden(name)
    dtype = d["a0"].dtype.type
    f64 = dtype is np.float64
    Nx, Ny = d["psi0"].shape
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, homogeneous_external_field=float(d["H"]))
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    gl = GLSolver(**kw)
    _set_state(gl, d, "psi0", "a0", "b0")
    gl.solve.cg(n_iter=25)
    E1 = np.array(gl.solve._cg.cg_energies)
    if f64:
        assert len(E1) == len(d["E1"]) and np.allclose(E1, d["E1"], rtol=1e-10)
        assert relerr(gl.vars.order_parameter, d["psi1"]) < 1e-9
        gl.solve.cg(n_iter=5)
        assert np.allclose(gl.solve._cg.cg_energies, d["E2"], rtol=1e-10)
        assert relerr(gl.vars.order_parameter, d["psi2"]) < 1e-9
    else:
        n = min(len(E1), len(d["E1"]))
        assert np.allclose(E1[:n], d["E1"][:n], rtol=2e-3)


@pytest.mark.parametrize("name", ["cg_f64_k2", "cg_f64_k2_tiled_eps"])
def test_cg_full_first_iterations(name):
    """Finite kappa: SciPy BFGS turns 1e-15 coefficient noise into ~1e-9 alpha noise after a few
    iterations (see tests/test_oracle_golden.py), so only the first iterations are tight."""
    from svirl_b200 import GLSolver
    d = load_golden(name)
    Nx, Ny = d["psi0"].shape
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, homogeneous_external_field=float(d["H"]),
              gl_parameter=float(d["kappa"]), normal_conductivity=10.0)
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    if bool(d["eps_is_field"]):
        kw["linear_coefficient"] = d["eps"]
    gl = GLSolver(**kw)
    _set_state(gl, d, "psi0", "a0", "b0")
    gl.solve.cg(n_iter=8)
    E = np.array(gl.solve._cg.cg_energies)
    assert np.allclose(E[:3], d["E1"][:3], rtol=1e-9)
    assert np.allclose(E, d["E1"][:len(E)], rtol=1e-3)
    assert np.all(np.diff(E) < 0)          # energy decreases monotonically


@pytest.mark.parametrize("kernel", [KERNELS[0], KERNELS[1], KERNELS[6], KERNELS[10], KERNELS[11]],
                         ids=["plain", "stream_k4_tma", "tile_k4", "tile_k4_plainA", "small"])
def test_cfg1_readme_1000_steps(kernel):
    """BASELINE configs[0]: 129^2, kappa 5, sigma 200, H 0.1, fp64, td(0.1, 1000): psi, a, b within
    1e-10, identical sweep counts, identical vortex count and positions."""
    from svirl_b200 import GLSolver
    d200, d1000 = load_golden("cfg1_td200"), load_golden("cfg1_td1000")
    gl = GLSolver(Lx=64, Ly=64, dx=0.5, dy=0.5, gl_parameter=5.0, normal_conductivity=200.0,
                  homogeneous_external_field=0.1, random_seed=1234)
    set_kernel(gl, kernel)
    assert np.array_equal(gl.vars.order_parameter, d200["psi0"])
    gl.solve.td(dt=0.1, Nt=200)
    td = gl.solve._td
    assert (td.sweeps_order_parameter, td.sweeps_vector_potential) == (int(d200["sweeps_psi"]), int(d200["sweeps_A"]))
    assert relerr(gl.vars.order_parameter, d200["psi1"]) < 1e-10
    gl.solve.td(dt=0.1, Nt=800)
    assert (td.sweeps_order_parameter, td.sweeps_vector_potential) == (int(d1000["sweeps_psi"]), int(d1000["sweeps_A"]))
    a, b = gl.vars.vector_potential
    assert relerr(gl.vars.order_parameter, d1000["psi1"]) < 1e-10
    assert relerr(a, d1000["a1"]) < 1e-10 and relerr(b, d1000["b1"]) < 1e-10
    assert abs(gl.observables.free_energy - d1000["obs_E"]) < 1e-10 * abs(d1000["obs_E"])
    vx, vy, vv = gl.vortex_detector.vortices
    assert vx.size == d1000["obs_vx"].size == 57
    assert np.array_equal(vv, d1000["obs_vv"])
    assert np.allclose(vx, d1000["obs_vx"], rtol=0, atol=1e-8) and np.allclose(vy, d1000["obs_vy"], rtol=0, atol=1e-8)


def test_reductions_like_reference_at_reduction():
    """tests/at_reduction.py of the reference: gsum / gsum_v against np.sum, atol 1e-10."""
    from svirl_b200 import GLSolver
    gl = GLSolver(Nx=16, Ny=16, dx=0.5, dy=0.5)
    rs = np.random.RandomState(5)
    for N in [1, 2, 31, 32, 33, 1000, 4097, 1234567]:
        a = rs.rand(N)
        assert abs(gl.par.red.test_sum(a, N, block_size=128) - a.sum()) < 1e-10 * max(N, 1)
        v = rs.rand(N, 5)
        out = gl.par.red.test_sum_v(v, N, 5, block_size=128)
        assert np.allclose(out, v.sum(axis=0), rtol=0, atol=1e-10 * max(N, 1))


def test_kernel_level_sweeps_match_oracle():
    """svl_td_psi_sweep / svl_td_a_sweep (the kernel-level entry points) against one oracle sweep."""
    import ctypes as C
    import glnumpy as O
    from svirl_b200 import GLSolver, _lib
    from svirl_b200.storage.arrays import DeviceArray
    d = load_golden("td_f64_k2_tiled_eps")
    m = d["meta"]
    gl = make_solver(m, d)
    g = O.Grid(m["Nx"], m["Ny"], m["dx"], m["dy"], np.float64)
    mt, eps = golden_inputs(d)
    psi0, a0, b0 = d["psi1"], d["a1"], d["b1"]
    gl.vars.order_parameter = psi0
    gl.vars.vector_potential = (a0, b0)
    par = gl.par
    rhs = gl.vars._psi.get_d_obj().copy()
    out = DeviceArray.zeros(par, _lib.NODE_C)
    r = C.c_double()
    _lib.call("svl_td_psi_sweep", par.ctx, 0.1, 0.0, gl.params.linear_coefficient_h().handle,
              gl.vars.vector_potential_h().handle, rhs.handle, gl.vars.order_parameter_h().handle, out.handle,
              0.0, 0, 1, C.byref(r))
    nxt, ro, _ = O.psi_sweep(g, 0.1, eps, mt, a0, b0, psi0, psi0)
    got = gl.unflatten_array(out.get())
    assert np.abs(got - nxt).max() < 1e-14 and abs(r.value - ro) < 1e-15
    ab = gl.vars.vector_potential_h()
    rhs_e, out_e = ab.copy(), DeviceArray.zeros(par, _lib.EDGE)
    _lib.call("svl_td_a_sweep", par.ctx, 0.1, 4.0, 0.1, 0.1, gl.vars.order_parameter_h().handle, ab.handle,
              rhs_e.handle, ab.handle, out_e.handle, 0.0, 0, 1, C.byref(r))
    an, bn, ro, _, _ = O.a_sweep(g, 0.1, 4.0, 0.1, 0.1, mt, psi0, a0, b0, a0, b0, a0, b0)
    flat = out_e.get()
    assert np.abs(gl.unflatten_a_array(flat[:gl.cfg.Na]) - an).max() < 1e-14
    assert np.abs(gl.unflatten_b_array(flat[gl.cfg.Na:]) - bn).max() < 1e-14
    assert abs(r.value - ro) < 1e-15


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stream_kernel_equals_plain_kernel_large_grid(dtype):
    """Size-independent property at a multi-strip, multi-segment size: the temporally blocked
    streaming kernel and the plain kernel follow the same trajectory (same sweep counts; values
    equal to rounding) on a 700 x 517 tiled grid with a disordered linear coefficient."""
    from svirl_b200 import GLSolver
    Nx, Ny = 700, 517
    rs = np.random.RandomState(3)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.15
    eps = (0.7 + 0.3 * rs.rand(Nx, Ny)).astype(dtype)
    out = []
    for kernel in (KERNELS[0], KERNELS[1], KERNELS[5], KERNELS[3], KERNELS[6], KERNELS[9]):
        gl = GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, dtype=dtype, homogeneous_external_field=0.1, random_seed=5,
                      material_tiling=mt, linear_coefficient=eps)
        set_kernel(gl, kernel)
        gl.solve.td(dt=0.1, Nt=12)
        out.append((gl.vars.order_parameter, gl.solve._td.sweeps_order_parameter))
        gl.par.close()
    tol = 1e-12 if dtype is np.float64 else 2e-5
    for psi, n in out[1:]:
        if dtype is np.float64:
            assert n == out[0][1]
        assert np.abs(psi - out[0][0]).max() < tol


@pytest.mark.parametrize("lang", [0.0, 0.05], ids=["quiet", "langevin"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_a_tile_kernel_equals_plain_kernel_large_grid(dtype, lang):
    """The pair-fusing A-sweep tile kernel (a_tile.cu) and the per-node A kernel follow the same
    finite-kappa trajectory on a multi-tile 700 x 517 grid with holes (same sweep counts in fp64,
    values equal to rounding), with and without Langevin noise."""
    from svirl_b200 import GLSolver
    Nx, Ny = 700, 517
    rs = np.random.RandomState(3)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.15
    out = []
    for ak in (0, 1):
        gl = GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, gl_parameter=2.0, normal_conductivity=10.0,
                      homogeneous_external_field=0.1, random_seed=5, material_tiling=mt,
                      order_parameter_Langevin_coefficient=lang, vector_potential_Langevin_coefficient=lang)
        gl.par.set_option("a_kernel", ak)
        gl.solve.td(dt=0.1, Nt=12)
        a, b = gl.vars.vector_potential
        out.append((gl.vars.order_parameter, a, b, gl.solve._td.sweeps_vector_potential, gl.par.stat("replays")))
        gl.par.close()
    tol = 1e-12 if dtype is np.float64 else 2e-5
    if dtype is np.float64:
        assert out[0][3] == out[1][3]
    for k in range(3):
        assert np.abs(out[0][k] - out[1][k]).max() < tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sincos_accuracy(dtype):
    """The library's own sincos (every link variable exp(-i d A) goes through it) against the host
    libm in extended precision: <= 2 ulp over the argument range the fast path covers, and exact
    hand-over to the slow path beyond it."""
    import ctypes as C
    from svirl_b200 import GLSolver, _lib
    gl = GLSolver(Nx=16, Ny=16, dx=0.5, dy=0.5, dtype=dtype)
    rs = np.random.RandomState(9)
    lim = 1.0e5 if dtype is np.float64 else 2.0e4
    x = np.concatenate([rs.uniform(-4, 4, 200000), rs.uniform(-300, 300, 200000), rs.uniform(-lim, lim, 200000),
                        np.arange(-64, 65) * (np.pi / 4), [0.0, 1e-300, -1e-9, 3 * lim, -7 * lim, 1e9]])
    x = x.astype(dtype).astype(np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.call("svl_debug_sincos", gl.par.ctx, x.size, pd(x), pd(s), pd(c))
    xl = x.astype(np.longdouble)
    es, ec = np.abs(s - np.sin(xl)).astype(np.float64), np.abs(c - np.cos(xl)).astype(np.float64)
    ulp = np.finfo(dtype).eps            # results are in [-1, 1]: absolute error in units of eps
    assert es.max() < 2.0 * ulp and ec.max() < 2.0 * ulp, (es.max() / ulp, ec.max() / ulp)


@pytest.mark.parametrize("case", ["f64_k2_holes_epsfield_ext", "f64_kinf_holes", "f32_k2_holes", "f32_kinf"])
def test_cg_fused_iteration_equals_kernel_composition(case):
    """The three-pass CG iteration (cg_fused.cu: Jacobians+PR sums / direction+coefficients /
    update+energy) against the composition of the single kernels (option cg_fused=0, the kernels
    checked one by one against the reference in test_kernels) on a multi-tile 300 x 270 grid:
    same energies and same state to rounding, including a second cg() call (quirk Q6)."""
    from svirl_b200 import GLSolver
    dtype = np.float64 if case.startswith("f64") else np.float32
    Nx, Ny = 300, 270
    rs = np.random.RandomState(11)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.1
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, homogeneous_external_field=0.1, random_seed=7,
              gl_parameter=np.inf if "kinf" in case else 2.0, normal_conductivity=10.0)
    if "holes" in case:
        kw["material_tiling"] = mt
    if "epsfield" in case:
        kw["linear_coefficient"] = (0.7 + 0.3 * rs.rand(Nx, Ny)).astype(dtype)
    res = []
    for fused in (0, 1, 2):        # 2 = two-pass iteration (cg_pipe.cu); with an external potential it falls back to 1
        gl = GLSolver(**kw)
        if "ext" in case:
            r2 = np.random.RandomState(3)
            gl.params.external_vector_potential = (0.05 * r2.rand(Nx - 1, Ny).astype(dtype), 0.05 * r2.rand(Nx, Ny - 1).astype(dtype))
        gl.par.set_option("cg_fused", fused)
        gl.solve.td(dt=0.1, Nt=5)
        gl.solve.cg(n_iter=4)
        E1 = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        gl.solve.cg(n_iter=3)
        E2 = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        a, b = gl.vars.vector_potential
        res.append((E1, E2, gl.vars.order_parameter, a, b))
        gl.par.close()
    f64 = dtype is np.float64
    # finite kappa goes through SciPy BFGS, which amplifies rounding differences of the coefficients
    # (the energies are sums with heavy cancellation: compare on the scale of the first one)
    rt = (1e-8 if "kinf" not in case else 1e-12) if f64 else 2e-4
    at = rt * np.abs(res[0][0]).max()
    tol = (1e-7 if "kinf" not in case else 1e-11) if f64 else 2e-3
    for r in res[1:]:
        assert np.allclose(res[0][0], r[0], rtol=rt, atol=at) and np.allclose(res[0][1], r[1], rtol=rt, atol=at)
        for k in (2, 3, 4):
            assert np.abs(res[0][k] - r[k]).max() < tol


def test_cg_line_search_rescue_on_large_grid_coefficients():
    """Coefficients measured at 16384^2 (per node, scaled back up): the reference's SciPy BFGS call
    runs away to |alpha| ~ 1e35 on them; the guarded line search detects that and returns the local
    minimum next to alpha = 0 (the one BFGS finds on the per-node polynomial)."""
    from svirl_b200 import GLSolver
    gl = GLSolver(Nx=16, Ny=16, dx=0.5, dy=0.5, gl_parameter=2.0)
    gl.solve._init_cg()
    cg = gl.solve._cg
    c = np.array([[-5.569412e-02, -5.033755e-03, 7.153037e-02, 2.296557e-04, -4.050537e-04],
                  [-2.623664e-02, -1.681635e-02, 2.434797e-02, 1.979796e-04, -2.142220e-04],
                  [2.284921e-02, -3.724204e-03, 5.584273e-03, 3.571101e-05, -4.655317e-05],
                  [1.439366e-02, 0, 0, 0, 0], [2.060112e-03, 0, 0, 0, 0]])
    cg._CG__c[:] = c * 16384.0 ** 2
    with np.errstate(all="ignore"):
        raw = cg._cg_alpha_min()
    assert not (np.all(np.isfinite(raw)) and np.max(np.abs(raw)) < 1e6)       # the reference call is unusable here
    a = cg._cg_alpha_min_guarded()
    assert cg.line_search_rescues == 1
    assert np.allclose(a, [0.42212742, 0.07723854], rtol=1e-4)
    gl.cfg.cg_line_search = "normalized"
    assert np.allclose(cg._cg_alpha_min_guarded(), [0.42212742, 0.07723854], rtol=1e-4) and cg.line_search_rescues == 1
    gl.cfg.cg_line_search = "reference"
    # well-scaled coefficients: the reference call is kept as is
    cg._CG__c[:] = c
    assert np.allclose(cg._cg_alpha_min_guarded(), cg._cg_alpha_min(), rtol=0, atol=0) and cg.line_search_rescues == 1


@pytest.mark.parametrize("name", ["td_f64_k2_fixed", "td_f32_kinf_fixed", "td_f64_k3_fixed_nolock"])
def test_td_fixed_vortices(name):
    """fixed vortices + phase lock through GLSolver against the unmodified reference
    (oracle/make_golden_fixed.py), quirks included: partial fold of the packed edge array, lock list of
    i-indices, drifting device copy of the irregular potential, stale host copies."""
    from svirl_b200 import GLSolver
    d = load_golden(name)
    m = d["meta"]
    f64 = m["dtype"] == "float64"
    kw = {k: v for k, v in m.items() if k not in ("Nt", "Nt2", "dtype")}
    kw["dtype"] = np.dtype(m["dtype"]).type
    kw["fixed_vortices"] = [[8.2, 15.1], [7.3, 11.0], [1, -1]]
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    gl = GLSolver(**kw)
    fv = gl.params.fixed_vortices
    assert np.array_equal(gl.vars.order_parameter, d["psi0"])
    vx, vy, vv = fv.fixed_vortices
    assert np.array_equal(vx, d["fvx"]) and np.array_equal(vy, d["fvy"]) and np.array_equal(vv, d["fvv"])
    ai, bi = fv.irregular_vector_potential
    t0 = 1e-13 if f64 else 1e-5
    assert np.abs(ai - d["ai0"]).max() < t0 and np.abs(bi - d["bi0"]).max() < t0
    lock = fv._phase_lock_ns.get_h().ravel() if fv._phase_lock_ns is not None else np.zeros(0, np.int32)
    assert np.array_equal(lock, d["lock_ns"])
    gl.solve.td(dt=0.1, Nt=m["Nt"])
    td = gl.solve._td
    tol = 1e-10 if f64 else 1e-4
    if f64:
        assert (td.sweeps_order_parameter, td.sweeps_vector_potential) == (int(d["sweeps_psi"]), int(d["sweeps_A"]))
    a, b = gl.vars.vector_potential
    assert relerr(gl.vars.order_parameter, d["psi1"]) < tol
    assert relerr(a, d["a1"]) < tol and relerr(b, d["b1"]) < tol
    assert np.abs(fv._vpi.get_d_obj().get() - d["vpi_dev1"]).max() < tol * max(np.abs(d["vpi_dev1"]).max(), 1.0)
    ai, bi = fv.irregular_vector_potential                       # host copy: unchanged, like the reference's
    assert np.abs(ai - d["ai0"]).max() < t0 and np.abs(bi - d["bi0"]).max() < t0
    assert abs(gl.observables.free_energy - d["obs_E"]) < (1e-9 if f64 else 2e-4) * max(abs(d["obs_E"]), 1.0)
    assert np.abs(fv.fixed_vortices_phase - d["phase"]).max() < (1e-11 if f64 else 1e-3)
    vx, vy, vv = gl.vortex_detector.vortices
    assert np.array_equal(vv, d["obs_vv"])
    if f64:
        assert np.allclose(vx, d["obs_vx"], rtol=0, atol=1e-8) and np.allclose(vy, d["obs_vy"], rtol=0, atol=1e-8)
    gl.solve.td(dt=0.1, Nt=m["Nt2"], eqn="order_parameter")
    assert int(td._random_t) == int(d["rand_t"])
    if f64:
        assert td.sweeps_order_parameter == int(d["sweeps_psi2"])
    assert relerr(gl.vars.order_parameter, d["psi2"]) < tol
    a, b = gl.vars.vector_potential                              # host copy stays at the stage-one values
    assert relerr(a, d["a2"]) < tol and relerr(b, d["b2"]) < tol
    assert np.abs(gl.vars._vp.get_d_obj().get() - d["vp_dev2"]).max() < tol * max(np.abs(d["vp_dev2"]).max(), 1.0)


@pytest.mark.parametrize("graphs", [0, 1], ids=["batched", "small"])
@pytest.mark.parametrize("shape", [(4, 4), (5, 9), (33, 31), (64, 65), (130, 7), (181, 181)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_small_and_ragged_grids_against_oracle(shape, dtype, graphs):
    """Edge cases of the geometry: the minimum grid (4x4), sizes below / across one tile and one warp,
    odd pitches: TDGL (finite kappa, random holes, eps field) and a CG iteration against the NumPy oracle."""
    import glnumpy as O
    from svirl_b200 import GLSolver
    Nx, Ny = shape
    rs = np.random.RandomState(Nx * 131 + Ny)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.2
    eps = (0.7 + 0.3 * rs.rand(Nx, Ny)).astype(dtype)
    gl = GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, gl_parameter=2.0, normal_conductivity=10.0,
                  homogeneous_external_field=0.1, random_seed=3, material_tiling=mt, linear_coefficient=eps)
    gl.par.set_option("graphs", graphs)
    g = O.Grid(Nx, Ny, 0.5, 0.4, dtype)
    psi0 = gl.vars.order_parameter
    a0, b0 = [x.copy() for x in gl.vars.vector_potential]
    gl.solve.td(dt=0.1, Nt=6)
    counts = []
    po, ao, bo, _ = O.td_run(g, 0.1, 6, eps, mt, 2.0, 10.0, 0.1, psi0, a0, b0, rand_t=3, counts=counts)
    f64 = dtype is np.float64
    tol = 1e-11 if f64 else 2e-4
    a1, b1 = gl.vars.vector_potential
    if f64:
        td = gl.solve._td
        assert (td.sweeps_order_parameter, td.sweeps_vector_potential) == (sum(c[0] for c in counts), sum(c[1] for c in counts))
    assert np.abs(gl.vars.order_parameter - po).max() < tol
    assert np.abs(a1 - ao).max() < tol and np.abs(b1 - bo).max() < tol
    k2 = dtype(dtype(2.0) ** 2)
    z = np.zeros_like
    Eo = O.free_energy(g, k2, eps, 0.1, mt, po, z(ao), z(bo), ao, bo)
    assert abs(gl.observables.free_energy - Eo) < (1e-10 if f64 else 2e-4) * max(abs(Eo), 1.0)
    gl.solve.cg(n_iter=1)
    _, _, _, Eo1, _ = O.cg_run(g, 1, 2.0, eps, 0.1, mt, po, z(ao), z(bo), ao, bo)
    assert abs(gl.solve._cg.cg_energies[0] - Eo1[0]) < (1e-8 if f64 else 1e-3) * max(abs(Eo1[0]), 1.0)


@pytest.mark.parametrize("case", ["f64_k2", "f32_kinf"])
def test_scale_driver_equals_glsolver_bitwise(case):
    """svirl_b200.scale.ScaleTD (slab-sized host buffers, fields generated row band by row band) follows
    the same trajectory as GLSolver BITWISE, and its GPU vortex count equals the detector's."""
    from svirl_b200 import GLSolver
    from svirl_b200.scale import ScaleTD
    dtype = np.float64 if case.startswith("f64") else np.float32
    kw = dict(Nx=300, Ny=270, dx=0.5, dy=0.5, dtype=dtype, homogeneous_external_field=0.1, random_seed=1234,
              gl_parameter=2.0 if "k2" in case else np.inf, normal_conductivity=10.0)
    gl = GLSolver(**kw)
    gl.solve.td(dt=0.1, Nt=25)
    psi, (a, b) = gl.vars.order_parameter, gl.vars.vector_potential
    sweeps = (gl.solve._td.sweeps_order_parameter, gl.solve._td.sweeps_vector_potential)
    vx, vy, vv = gl.vortex_detector.vortices
    gl.par.close()
    st = ScaleTD(band_rows=64, **kw)
    st.td(0.1, 25)
    assert (st.sweeps[0], st.sweeps[1]) == sweeps
    assert np.array_equal(st.psi_rows(0, 270), psi)
    assert np.array_equal(st.a_rows(0, 270), a) and np.array_equal(st.b_rows(0, 269), b)
    assert np.array_equal(st.psi_rows(100, 131), psi[:, 100:131])
    import glnumpy as O
    v = O.winding(O.Grid(300, 270, 0.5, 0.5, dtype), 0.1, psi, a, b)        # the detector's winding test on the host
    ok = (np.abs(v) > 0.5) & (np.abs(v - np.round(v)) < 0.1)
    assert st.vortex_count() == (int(np.sum(ok & (v > 0))), int(np.sum(ok & (v < 0))))
    assert vv.size <= int(ok.sum())                                            # triangulation may only reject
    st.close()
