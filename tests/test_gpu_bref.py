"""GPU: three-way parity  reference fixtures (reference on the CPU shim)  <->  B-ref (the unmodified reference
package with its own CUDA kernels on this GPU, baseline/bref.py)  <->  svirl_b200, including BASELINE-sized grids
(cfg2 2048^2 fp32 tiled, cfg3 8192^2 fp64 kappa=2 with a disordered linear coefficient).

Tolerances (north_star): fp64 1e-10 relative, fp32 1e-4; fp64 Jacobi sweep counts must be identical."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, has_cuda

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import bref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device"),
              pytest.mark.skipif(not bref.available(), reason="baseline/_ref not installed")]


def relerr(x, ref):
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300))


def hole_tiling(Nx, Ny, dx=0.5, dy=0.5):
    x = (np.arange(Nx - 1) + 0.5) * dx
    y = (np.arange(Ny - 1) + 0.5) * dy
    return ~((((np.mod(x, 16.0) - 8.0) ** 2)[:, None] + ((np.mod(y, 16.0) - 8.0) ** 2)[None, :]) < 4.0)


def run_both(kw, Nt, dt=0.1, cg_iters=0):
    """The same constructor keywords through the reference (B-ref) and through svirl_b200."""
    from svirl_b200 import GLSolver
    ref = bref.make_solver(**kw)
    c = bref.launch_counts()
    c.clear()
    ref.solve.td(dt=dt, Nt=Nt)
    r = dict(zip(("psi", "a", "b"), bref.fields(ref)))
    r["sweeps"] = (c.get("iterate_order_parameter_jacobi_step", 0), c.get("iterate_vector_potential_jacobi_step", 0))
    gl = GLSolver(**kw)
    gl.solve.td(dt=dt, Nt=Nt)
    o = dict(psi=gl.vars.order_parameter, a=gl.vars.vector_potential[0].copy(), b=gl.vars.vector_potential[1].copy())
    o["sweeps"] = (gl.solve._td.sweeps_order_parameter, gl.solve._td.sweeps_vector_potential)
    if cg_iters:
        ref.solve.cg(n_iter=cg_iters)
        r["E"] = np.array(ref.solve._cg.cg_energies, dtype=np.float64)
        r["psi_cg"] = bref.fields(ref)[0]
        gl.solve.cg(n_iter=cg_iters)
        o["E"] = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
        o["psi_cg"] = gl.vars.order_parameter
    r["E_td"], o["E_td"] = float(ref.observables.free_energy), float(gl.observables.free_energy)
    gl.par.close()
    del ref
    return r, o


def test_bref_reproduces_reference_fixture():
    """The pyCUDA stand-in is faithful: the reference on the GPU reproduces the fixture the reference produced on
    the CPU shim (td_f64_k5: 37x29, kappa 5, 60 steps) to rounding, with the same sweep counts."""
    d = load_golden("td_f64_k5")
    m = d["meta"]
    kw = {k: v for k, v in m.items() if k not in ("Nt", "dtype")}
    ref = bref.make_solver(dtype=np.float64, **kw)
    psi0 = bref.fields(ref)[0]
    assert np.array_equal(psi0, d["psi0"])
    c = bref.launch_counts()
    c.clear()
    ref.solve.td(dt=0.1, Nt=m["Nt"])
    psi, a, b = bref.fields(ref)
    assert (c["iterate_order_parameter_jacobi_step"], c["iterate_vector_potential_jacobi_step"]) == \
        (int(d["sweeps_psi"]), int(d["sweeps_A"]))
    assert relerr(psi, d["psi1"]) < 1e-11 and relerr(a, d["a1"]) < 1e-11 and relerr(b, d["b1"]) < 1e-11
    vx, vy, vv = ref.vortex_detector.vortices
    assert np.array_equal(vv, d["obs_vv"]) and np.allclose(vx, d["obs_vx"], atol=1e-8, rtol=0)


@pytest.mark.parametrize("case", ["f64_k2", "f64_kinf", "f32_kinf"])
def test_td_against_reference_on_gpu(case):
    Nx, Ny = 300, 270
    rs = np.random.RandomState(11)
    kw = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.5, homogeneous_external_field=0.1, random_seed=1234,
              material_tiling=rs.rand(Nx - 1, Ny - 1) > 0.1, dtype=np.float32 if "f32" in case else np.float64)
    if "k2" in case:
        kw.update(gl_parameter=2.0, normal_conductivity=10.0)
    r, o = run_both(kw, 25, cg_iters=4 if case != "f32_kinf" else 0)
    f64 = "f64" in case
    tol = 1e-10 if f64 else 1e-4
    if f64:
        assert r["sweeps"] == o["sweeps"]
    assert relerr(o["psi"], r["psi"]) < tol and relerr(o["a"], r["a"]) < tol and relerr(o["b"], r["b"]) < tol
    assert abs(o["E_td"] - r["E_td"]) < tol * abs(r["E_td"])
    if "E" in r:
        assert np.allclose(o["E"], r["E"], rtol=1e-9 if case == "f64_kinf" else 1e-7)


def test_cfg2_full_size_against_reference_on_gpu():
    """BASELINE configs[1] at full size: 2048^2 fp32, kappa=inf, hole-lattice tiling, 25 steps."""
    kw = dict(Nx=2048, Ny=2048, dx=0.5, dy=0.5, dtype=np.float32, homogeneous_external_field=0.1, random_seed=1234,
              material_tiling=hole_tiling(2048, 2048))
    r, o = run_both(kw, 25)
    assert abs(o["sweeps"][0] - r["sweeps"][0]) <= 0.05 * r["sweeps"][0]
    assert relerr(o["psi"], r["psi"]) < 1e-4


def test_cfg3_full_size_against_reference_on_gpu():
    """BASELINE configs[2] at full size: 8192^2 fp64, kappa=2, sigma=10, disordered linear coefficient; 2 steps
    (psi AND A kernels, reference launch pattern td.py:164-202, 274-311): <= 1e-10, identical sweep counts."""
    N = 8192
    eps = 0.7 + 0.3 * np.random.RandomState(4321).rand(N, N)
    kw = dict(Nx=N, Ny=N, dx=0.5, dy=0.5, dtype=np.float64, gl_parameter=2.0, normal_conductivity=10.0,
              homogeneous_external_field=0.1, random_seed=1234, linear_coefficient=eps)
    r, o = run_both(kw, 2)
    assert r["sweeps"] == o["sweeps"], (r["sweeps"], o["sweeps"])
    assert relerr(o["psi"], r["psi"]) < 1e-10
    assert relerr(o["a"], r["a"]) < 1e-10 and relerr(o["b"], r["b"]) < 1e-10
    assert abs(o["E_td"] - r["E_td"]) < 1e-10 * abs(r["E_td"])


def test_cfg4_cg_follows_reference_until_the_reference_diverges():
    """BASELINE configs[3] at 8192^2 (kappa 2, fp64, 20 TDGL steps, then CG).  The fixture holds what the UNMODIFIED
    reference did on a B200 (tools/cfg4_adjudicate.py, profiles/r02_cfg4_adjudicate_8192.json): its 4th SciPy BFGS
    line search returns alpha ~ -1e142 and the state is lost (E = inf, then NaN).  This library must reproduce the
    reference's energies while the reference is finite, and -- with the rescue of solvers/cg.py -- keep descending."""
    import json
    from svirl_b200 import GLSolver
    with open(os.path.join(ROOT, "tests", "golden", "cfg4_reference_cg_8192.json")) as f:
        fx = json.load(f)
    Er = np.array(fx["reference_energies"])
    nfin = int(np.argmin(np.isfinite(Er))) if not np.isfinite(Er).all() else len(Er)
    assert nfin == 3                                     # the reference survives three iterations
    c = fx["config"]
    gl = GLSolver(Nx=c["Nx"], Ny=c["Ny"], dx=c["dx"], dy=c["dy"], dtype=np.float64, gl_parameter=c["gl_parameter"],
                  normal_conductivity=c["normal_conductivity"], homogeneous_external_field=c["homogeneous_external_field"],
                  random_seed=c["random_seed"])
    gl.solve.td(dt=0.1, Nt=fx["td_steps"])
    gl.solve._init_cg()
    gl.solve._cg._CG__convergence_rtol = -1.0
    gl.solve.cg(n_iter=8)
    E = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
    assert np.allclose(E[:nfin], Er[:nfin], rtol=1e-8)
    assert np.all(np.isfinite(E)) and np.all(np.diff(E) < 0)
    assert gl.solve._cg.line_search_rescues >= 1
    gl.par.close()
