"""GPU: parity cases added in round 2 (VERDICT r01 "next" list and ADVICE r01).

* vortex detector BITWISE against the reference's own end states (north_star: "vortex count and positions
  must match exactly"), scalar and vectorised host path;
* the Jacobi diagonal vanishing on inactive nodes (dt * eps == 1), hole tiling and ragged grid;
* BASELINE-sized grids: see tests/test_gpu_fullsize.py."""
import numpy as np
import pytest

from conftest import load_golden, golden_inputs, has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

VORTEX_CASES = ["cfg1_td200", "cfg1_td1000", "td_f64_k5", "td_f64_kinf", "cg_f64_k2", "cg_f64_kinf", "td_f32_k3_langevin"]


def _solver_for_state(name):
    """A solver with the fixture's parameters holding the REFERENCE's end state (the one its detector saw)."""
    from svirl_b200 import GLSolver
    d = load_golden(name)
    if name.startswith("cfg1"):
        m = load_golden("cfg1_td200")["meta"]
    elif name.startswith("cg_"):
        m = None
    else:
        m = d["meta"]
    if m is not None:
        kw = {k: v for k, v in m.items() if k not in ("Nt", "dtype")}
        kw["dtype"] = np.dtype(m["dtype"]).type
    else:
        Nx, Ny = d["psi1"].shape
        kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=d["a1"].dtype.type, homogeneous_external_field=float(d["H"]),
                  gl_parameter=float(d["kappa"]))
    if "mt" in d:
        kw["material_tiling"] = d["mt"]
    gl = GLSolver(**kw)
    k = "2" if "psi2" in d else "1"        # CG fixtures store the observables after the second cg() call
    gl.vars.order_parameter = d["psi" + k]
    gl.vars.vector_potential = (d["a" + k], d["b" + k])
    return gl, d


@pytest.mark.parametrize("vector", [False, True], ids=["scalar", "vectorised"])
@pytest.mark.parametrize("name", VORTEX_CASES)
def test_vortices_bitwise_on_reference_state(name, vector, monkeypatch):
    """The reference's end state goes through svl_vortex_candidates (GPU) + the host re-test: x, y and
    vorticity equal the reference detector's output bit for bit, on both host paths."""
    import svirl_b200.observables.vortex_detector as vd
    gl, d = _solver_for_state(name)
    monkeypatch.setattr(vd, "VECTOR_THRESHOLD", -1 if vector else 1 << 30)
    vx, vy, vv = gl.vortex_detector.vortices
    assert vx.dtype == d["obs_vx"].dtype
    assert np.array_equal(vx, d["obs_vx"]) and np.array_equal(vy, d["obs_vy"]) and np.array_equal(vv, d["obs_vv"])
    gl.par.close()


@pytest.mark.parametrize("shape", [(96, 80), (131, 71)], ids=["holes", "ragged"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", [(0, 1), (2, 4), (2, 1), (1, 4), (-1, 0)], ids=["plain", "tile_k4", "tile_k1", "stream_k4", "small"])
def test_jacobi_diagonal_vanishes_on_inactive_nodes(shape, dtype, kernel):
    """dt = 1.0 with linear coefficient 1.0: on inactive nodes D = 1 + dt*(0 - eps + 0) = 0.  The reference
    writes psi = 0 there (td.h:117); a branch-free 1/D must not seed NaNs."""
    import glnumpy as O
    from svirl_b200 import GLSolver
    Nx, Ny = shape
    rs = np.random.RandomState(7)
    if shape == (96, 80):
        x = (np.arange(Nx - 1) + 0.5)
        y = (np.arange(Ny - 1) + 0.5)
        mt = ~(((np.mod(x, 16.0) - 8.0) ** 2)[:, None] + ((np.mod(y, 16.0) - 8.0) ** 2)[None, :] < 9.0)
    else:
        mt = rs.rand(Nx - 1, Ny - 1) > 0.3
    gl = GLSolver(Nx=Nx, Ny=Ny, dx=1.0, dy=1.0, dtype=dtype, homogeneous_external_field=0.05, random_seed=3,
                  material_tiling=mt, linear_coefficient=1.0)
    gl.par.set_option("graphs", 1 if kernel[0] < 0 else 0)
    if kernel[0] >= 0:
        gl.par.set_option("psi_kernel", kernel[0])
        gl.par.set_option("psi_k", kernel[1])
    g = O.Grid(Nx, Ny, 1.0, 1.0, dtype)
    psi0 = gl.vars.order_parameter
    a0, b0 = [v.copy() for v in gl.vars.vector_potential]
    gl.solve.td(dt=1.0, Nt=4)
    counts = []
    po, _, _, _ = O.td_run(g, 1.0, 4, 1.0, mt, np.inf, 1.0, 0.05, psi0, a0, b0, rand_t=3, counts=counts)
    psi = gl.vars.order_parameter
    assert np.all(np.isfinite(psi.real)) and np.all(np.isfinite(psi.imag))
    f64 = dtype is np.float64
    if f64:
        assert gl.solve._td.sweeps_order_parameter == sum(c[0] for c in counts)
    assert np.abs(psi - po).max() < (1e-11 if f64 else 2e-4)
    # inactive nodes are exactly zero, as in the reference
    act = np.zeros((Nx, Ny), dtype=bool)
    act[:-1, :-1] |= mt; act[1:, :-1] |= mt; act[:-1, 1:] |= mt; act[1:, 1:] |= mt
    assert np.all(psi[~act] == 0)
    gl.par.close()


@pytest.mark.parametrize("shape", [(300, 270), (517, 700), (70, 40), (249, 65), (8, 5)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("case", ["f64_k2_eps", "f64_kinf", "f32_k2", "f32_kinf"])
def test_cg_two_pass_iteration_equals_kernel_composition(case, shape):
    """The two-pass CG iteration (cg_pipe.cu: update + energy + next Jacobians + PR sums / direction + coefficients)
    against the composition of the single kernels (option cg_fused = 0; those are checked one by one against the
    reference in test_kernels), on strip-boundary and ragged sizes: same energies and state to rounding over two
    cg() calls (quirk Q6: beta and, for finite kappa, the directions persist), the second one ending by convergence."""
    from svirl_b200 import GLSolver
    dtype = np.float64 if case.startswith("f64") else np.float32
    Nx, Ny = shape
    rs = np.random.RandomState(Nx + Ny)
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=dtype, homogeneous_external_field=0.1, random_seed=7,
              gl_parameter=np.inf if "kinf" in case else 2.0, normal_conductivity=10.0,
              material_tiling=rs.rand(Nx - 1, Ny - 1) > 0.1)
    if "eps" in case:
        kw["linear_coefficient"] = (0.7 + 0.3 * rs.rand(Nx, Ny)).astype(dtype)
    res = []
    for fused in (0, 2):
        gl = GLSolver(**kw)
        gl.par.set_option("cg_fused", fused)
        gl.solve.td(dt=0.1, Nt=5)
        gl.solve.cg(n_iter=4)
        E1 = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        gl.cfg.convergence_rtol = 0.5
        gl.solve._cg._CG__convergence_rtol = 0.5          # stops after the second iteration (i = 1)
        gl.solve.cg(n_iter=5)
        E2 = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        gl.solve._cg._CG__convergence_rtol = -1.0
        gl.solve.cg(n_iter=2)
        E3 = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        a, b = gl.vars.vector_potential
        res.append((E1, E2, E3, gl.vars.order_parameter, a.copy(), b.copy()))
        gl.par.close()
    f64 = dtype is np.float64
    assert len(res[0][1]) == len(res[1][1]) < 5             # both stopped by the convergence test, at the same iteration
    rt = (1e-8 if "kinf" not in case else 1e-11) if f64 else 3e-4
    at = rt * np.abs(res[0][0]).max()
    for k in (0, 1, 2):
        assert np.allclose(res[0][k], res[1][k], rtol=rt, atol=at), (k, res[0][k], res[1][k])
    tol = (1e-7 if "kinf" not in case else 1e-10) if f64 else 3e-3
    for k in (3, 4, 5):
        assert np.abs(res[0][k] - res[1][k]).max() < tol


@pytest.mark.parametrize("case", ["f64_k2", "f32_kinf"])
def test_scale_driver_checkpoint_restart_is_bitwise(case, tmp_path):
    """a slab checkpoint (.npz) restored into a fresh driver continues the run bit for bit
    (fields, Langevin counter, sweep counts)."""
    from svirl_b200.scale import ScaleTD
    dtype = np.float64 if case.startswith("f64") else np.float32
    kw = dict(Nx=300, Ny=270, dx=0.5, dy=0.5, dtype=dtype, homogeneous_external_field=0.1, random_seed=1234,
              gl_parameter=2.0 if "k2" in case else np.inf, normal_conductivity=10.0)
    st = ScaleTD(band_rows=64, **kw)
    st.td(0.1, 10)
    st.save_checkpoint(str(tmp_path / "ck"))
    st.td(0.1, 10)
    want = (st.psi_rows(0, 270), st.a_rows(0, 270), st.b_rows(0, 269), st.sweeps[0], st.sweeps[1], st.rand_t.value)
    st.close()
    st2 = ScaleTD(band_rows=64, **dict(kw, random_seed=99))           # different initial state: everything comes from the file
    st2.load_checkpoint(str(tmp_path / "ck"))
    st2.td(0.1, 10)
    assert (st2.sweeps[0], st2.sweeps[1], st2.rand_t.value) == want[3:]
    assert np.array_equal(st2.psi_rows(0, 270), want[0])
    assert np.array_equal(st2.a_rows(0, 270), want[1]) and np.array_equal(st2.b_rows(0, 269), want[2])
    st2.close()


def test_host_step_pipeline_equals_sequential_steps():
    """Three solver instances driven by three host threads (upload, one step, download per instance-step,
    svirl_b200/parallel/pipeline.py) end in the states the same instances reach one after the other, bit for bit:
    contexts share nothing (per-context streams, scratch, tensor-map cache)."""
    import ctypes as C
    from svirl_b200 import GLSolver, _lib
    from svirl_b200.parallel.pipeline import HostStepPipeline
    Nx, Ny, nst = 300, 270, 4
    mt = np.ones((Nx - 1, Ny - 1), dtype=bool)
    mt[40:60, 30:50] = False

    def make(seed):
        return GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, dtype=np.float32, gl_parameter=np.inf,
                        homogeneous_external_field=0.1, random_seed=seed, material_tiling=mt)

    seeds = [3, 4, 5]
    want = []
    for s in seeds:                                   # sequential: psi stays on the device
        gl = make(s)
        gl.solve.td(dt=0.1, Nt=nst)
        want.append(gl.flatten_array(gl.vars.order_parameter).copy())
        gl.par.close()
    sols = [make(s) for s in seeds]
    hin = [np.ascontiguousarray(g.flatten_array(g.vars.order_parameter)) for g in sols]
    hout = [np.empty_like(h) for h in hin]
    secs = HostStepPipeline(sols).run(hin, hout, nst, dt=0.1)
    assert secs > 0
    for k, g in enumerate(sols):
        last = hout[k] if nst % 2 == 1 else hin[k]    # buffers swap roles every step
        assert np.array_equal(last, want[k]), k
        assert np.array_equal(g.flatten_array(g.vars.order_parameter), want[k]), k
        g.par.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pipelined_solves_equal_plain_solves_bitwise(dtype):
    """kappa = inf time stepping with the next step's first launch pre-issued behind the device-side stop rule
    (option pipeline, svirl_b200/csrc/td.cu) gives the same psi bit for bit and the same sweep counts as the plain
    driver, with and without Langevin noise; the gate must open at least once for the test to mean anything."""
    from svirl_b200 import GLSolver
    Nx, Ny = 300, 270
    mt = np.ones((Nx - 1, Ny - 1), dtype=bool)
    mt[100:140, 60:90] = False
    out = {}
    for lang in (0.0, 0.05):
        for pipe in (0, 1):
            gl = GLSolver(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, dtype=dtype, gl_parameter=np.inf, homogeneous_external_field=0.1,
                          random_seed=7, material_tiling=mt, order_parameter_Langevin_coefficient=lang)
            gl.par.set_option("pipeline", pipe)
            gl.solve.td(dt=0.1, Nt=6)
            gl.solve.td(dt=0.1, Nt=30)
            out[pipe] = (gl.flatten_array(gl.vars.order_parameter).copy(), gl.solve._td.sweeps_order_parameter,
                         gl.par.stat("spec_hit"), gl.par.stat("spec_miss"))
            gl.par.close()
        assert out[0][1] == out[1][1], (lang, out[0][1], out[1][1])
        assert np.array_equal(out[0][0], out[1][0]), lang
        assert out[0][2] == 0 and out[1][2] > 0, (lang, out[0][2:], out[1][2:])


@pytest.mark.parametrize("shape", [(300, 270), (131, 71)], ids=["multi_tile", "ragged"])
@pytest.mark.parametrize("epsfield", [False, True], ids=["eps_scalar", "eps_field"])
def test_patch_kernel_equals_column_kernel_bitwise(shape, epsfield):
    """k_psi_patch (2 x 4 nodes per thread, option psi_patch = 1, default) performs the column kernel's 16-FMA chain per
    node in the same order: fp32 psi and sweep counts are bit-identical with psi_patch = 0, with a spatially varying
    linear coefficient, holes, Langevin noise and grids that end inside a tile."""
    from svirl_b200 import GLSolver
    Nx, Ny = shape
    rs = np.random.RandomState(11)
    mt = rs.rand(Nx - 1, Ny - 1) > 0.08
    out = []
    for patch in (1, 0):
        kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, dtype=np.float32, gl_parameter=np.inf, homogeneous_external_field=0.15,
                  random_seed=9, material_tiling=mt, order_parameter_Langevin_coefficient=0.03)
        if epsfield:
            kw["linear_coefficient"] = (0.6 + 0.4 * np.random.RandomState(5).rand(Nx, Ny)).astype(np.float32)
        gl = GLSolver(**kw)
        gl.par.set_option("graphs", 0)
        gl.par.set_option("psi_patch", patch)
        gl.solve.td(dt=0.1, Nt=7)
        out.append((gl.flatten_array(gl.vars.order_parameter).copy(), gl.solve._td.sweeps_order_parameter))
        gl.par.close()
    assert out[0][1] == out[1][1]
    assert np.isfinite(out[0][0]).all() and np.array_equal(out[0][0], out[1][0])
