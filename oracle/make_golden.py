"""Generate tests/golden/*.npz by running the UNMODIFIED reference on the CPU -- TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # needs /root/reference (build container only)

Each fixture stores the reference's inputs and outputs for one seeded case; the
kernel-launch counts (= Jacobi sweep counts) come from the fake pyCUDA's launcher.
tests/test_oracle_golden.py checks oracle/glnumpy.py against them on the CPU and
tests/test_gpu_parity.py checks the CUDA path against them on the B200.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refrun  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def mt_holes(x, y):
    """Square lattice (period 8) of circular holes, radius^2 2.5 (scaled-down cfg2 tiling)."""
    return ~(((np.mod(x, 8) - 4) ** 2 + (np.mod(y, 8) - 4) ** 2) < 2.5)


def eps_field(Nx, Ny):
    return 0.7 + 0.3 * np.random.RandomState(4321).rand(Nx, Ny)


def state(gl):
    a, b = gl.vars.vector_potential
    return gl.vars.order_parameter.copy(), a.copy(), b.copy()


def observables(gl, d, tag):
    d[tag + "E"] = gl.observables.free_energy
    d[tag + "B"] = gl.observables.magnetic_field
    jx, jy = gl.observables.supercurrent_density
    d[tag + "jsx"], d[tag + "jsy"] = jx, jy
    if gl.params.solveA:
        jx, jy = gl.observables.current_density
        d[tag + "jx"], d[tag + "jy"] = jx, jy
    vx, vy, vv = gl.vortex_detector.vortices
    d[tag + "vx"], d[tag + "vy"], d[tag + "vv"] = vx, vy, vv


def td_case(svirl, name, Nt, **kw):
    refrun.launch_counts().clear()
    gl = svirl.GLSolver(**kw)
    d = {}
    d["psi0"], d["a0"], d["b0"] = state(gl)
    if kw.get("material_tiling") is not None:
        d["mt"] = gl.mesh.material_tiling
    if not np.isscalar(kw.get("linear_coefficient", 1.0)):
        d["eps"] = gl.params.linear_coefficient
    ae, be = gl.params.external_irregular_vector_potential
    d["ae"], d["be"] = ae.copy(), be.copy()
    gl.solve.td(dt=0.1, Nt=Nt)
    c = refrun.launch_counts()
    d["sweeps_psi"] = c.get("iterate_order_parameter_jacobi_step", 0)
    d["sweeps_A"] = c.get("iterate_vector_potential_jacobi_step", 0)
    d["psi1"], d["a1"], d["b1"] = state(gl)
    d["rand_t"] = int(gl.solve._td._random_t)
    observables(gl, d, "obs_")
    meta = {k: v for k, v in kw.items() if np.isscalar(v) and not callable(v) and k != "dtype"}
    meta["dtype"] = np.dtype(kw.get("dtype", np.float64)).name
    meta["Nt"] = Nt
    d["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "sweeps", d["sweeps_psi"], d["sweeps_A"], "vortices", d["obs_vx"].size, "E", d["obs_E"])
    return gl, d


def kernel_case(svirl, name, gl, seed=7):
    """Per-kernel outputs on the current state of ``gl`` with random directions."""
    from svirl.storage import GArray
    cfg = gl.cfg
    Nx, Ny = int(cfg.Nx), int(cfg.Ny)
    d = {}
    d["psi"], d["a"], d["b"] = state(gl)
    ae, be = gl.params.external_irregular_vector_potential
    d["ae"], d["be"] = ae.copy(), be.copy()
    if gl.mesh.have_material_tiling():
        d["mt"] = gl.mesh.material_tiling
    d["eps"] = gl.params.linear_coefficient
    d["eps_is_field"] = gl.params._epsilon.size != 1
    d["kappa2"] = gl.params.gl_parameter_squared_h()
    d["H"] = gl.params.homogeneous_external_field
    rs = np.random.RandomState(seed)
    dpsi_h = (rs.rand(Nx, Ny) - 0.5 + 1j * (rs.rand(Nx, Ny) - 0.5)).astype(cfg.dtype_complex)
    da_h = ((rs.rand(Nx - 1, Ny) - 0.5) * 0.1).astype(cfg.dtype)
    db_h = ((rs.rand(Nx, Ny - 1) - 0.5) * 0.1).astype(cfg.dtype)
    d["dpsi"], d["da"], d["db"] = dpsi_h, da_h, db_h
    gl.solve._init_cg()
    cg = gl.solve._cg
    d["E"] = gl.observables.free_energy
    d["jac_psi"] = gl.unflatten_array(cg._free_energy_jacobian_psi.get())
    dpsi = GArray(like=dpsi_h)
    if gl.params.solveA:
        jA = cg._free_energy_jacobian_A.get()
        d["jac_a"] = gl.unflatten_a_array(jA[:cfg.Na]).copy()
        d["jac_b"] = gl.unflatten_b_array(jA[cfg.Na:]).copy()
        dab = GArray(shape=[da_h.shape, db_h.shape], dtype=cfg.dtype)
        dab.set_vec_h(da_h, db_h)
        dab.sync()
        # before coef_psi: that call broadcasts its 5 sums into every row of the shared
        # 5x5 host matrix (cg.py:221 `self.__c[:] = gCr[0:5]`) and would leave stale rows 3-4
        d["coef17"] = np.array(cg._free_energy_conjgrad_coef(dpsi.get_d_obj(), dab.get_d_obj())).copy()
    c5 = np.array(cg._free_energy_conjgrad_coef_psi(dpsi.get_d_obj()))
    d["coef5"] = c5[0, :].copy() if c5.ndim == 2 else c5.copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "E", d["E"])


def cg_case(svirl, name, gl, n1, n2):
    d = {}
    d["psi0"], d["a0"], d["b0"] = state(gl)
    ae, be = gl.params.external_irregular_vector_potential
    d["ae"], d["be"] = ae.copy(), be.copy()
    if gl.mesh.have_material_tiling():
        d["mt"] = gl.mesh.material_tiling
    d["eps"] = gl.params.linear_coefficient
    d["eps_is_field"] = gl.params._epsilon.size != 1
    d["kappa"] = gl.params.gl_parameter
    d["H"] = gl.params.homogeneous_external_field
    gl.solve._init_cg()
    cg = gl.solve._cg
    alphas = []
    if gl.params.solveA:
        orig = cg._cg_alpha_min

        def wrap(*a, **k):
            r = orig(*a, **k)
            alphas.append(np.array(r))
            return r
        cg._cg_alpha_min = wrap
    else:
        orig = cg._cg_alpha_psi_min

        def wrap(*a, **k):
            r = orig(*a, **k)
            alphas.append(np.array([r]))
            return r
        cg._cg_alpha_psi_min = wrap
    gl.solve.cg(n_iter=n1)
    d["E1"] = np.array(cg.cg_energies)
    d["psi1"], d["a1"], d["b1"] = state(gl)
    gl.solve.cg(n_iter=n2)                       # quirk Q6: state persists across calls
    d["E2"] = np.array(cg.cg_energies)
    d["psi2"], d["a2"], d["b2"] = state(gl)
    d["alphas"] = np.array(alphas)
    observables(gl, d, "obs_")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "iters", len(d["E1"]), len(d["E2"]), "E", d["E2"][-1])


def main():
    svirl = refrun.import_reference()
    os.makedirs(OUT, exist_ok=True)
    Nx, Ny = 37, 29
    base = dict(Nx=Nx, Ny=Ny, dx=0.5, dy=0.4, homogeneous_external_field=0.1, random_seed=1234)

    # --- TD trajectories
    td_case(svirl, "td_f64_k5", 60, gl_parameter=5.0, normal_conductivity=200.0, **base)
    gl, _ = td_case(svirl, "td_f64_k2_tiled_eps", 150, gl_parameter=2.0, normal_conductivity=10.0,
                    material_tiling=mt_holes, linear_coefficient=eps_field(Nx, Ny), **base)
    cg_case(svirl, "cg_f64_k2_tiled_eps", gl, 8, 3)
    gl, _ = td_case(svirl, "td_f32_kinf_tiled", 60, dtype=np.float32, material_tiling=mt_holes, **base)
    cg_case(svirl, "cg_f32_kinf_tiled", gl, 25, 5)
    td_case(svirl, "td_f32_k2_tiled_eps", 60, dtype=np.float32, gl_parameter=2.0, normal_conductivity=10.0,
            material_tiling=mt_holes, linear_coefficient=eps_field(Nx, Ny), **base)
    gl, _ = td_case(svirl, "td_f64_kinf", 60, **base)
    cg_case(svirl, "cg_f64_kinf", gl, 25, 5)
    td_case(svirl, "td_f64_k3_langevin", 20, gl_parameter=3.0, normal_conductivity=200.0, material_tiling=mt_holes,
            linear_coefficient=eps_field(Nx, Ny), order_parameter_Langevin_coefficient=0.01,
            vector_potential_Langevin_coefficient=0.002, **base)
    td_case(svirl, "td_f32_k3_langevin", 20, dtype=np.float32, gl_parameter=3.0, normal_conductivity=200.0,
            material_tiling=mt_holes, order_parameter_Langevin_coefficient=0.01,
            vector_potential_Langevin_coefficient=0.002, **base)

    # --- per-kernel cases (external field + eps field + tiling)
    for dt_, nm in ((np.float64, "f64"), (np.float32, "f32")):
        gl = svirl.GLSolver(gl_parameter=3.0, normal_conductivity=200.0, dtype=dt_, material_tiling=mt_holes,
                            linear_coefficient=eps_field(Nx, Ny), external_field=0.05, **base)
        gl.solve.td(dt=0.1, Nt=20)
        kernel_case(svirl, "kernels_%s_k3_ext" % nm, gl)
        gl = svirl.GLSolver(dtype=dt_, external_field=0.05, **base)
        gl.solve.td(dt=0.1, Nt=20)
        kernel_case(svirl, "kernels_%s_kinf" % nm, gl)
    gl = svirl.GLSolver(gl_parameter=2.0, normal_conductivity=10.0, **base)
    gl.solve.td(dt=0.1, Nt=20)
    cg_case(svirl, "cg_f64_k2", gl, 8, 3)

    # --- cfg1 (README): 129^2, kappa 5, sigma 200, H 0.1, fp64, seed 1234
    cfg1 = dict(Lx=64, Ly=64, dx=0.5, dy=0.5, gl_parameter=5.0, normal_conductivity=200.0,
                homogeneous_external_field=0.1, random_seed=1234)
    gl, d200 = td_case(svirl, "cfg1_td200", 200, **cfg1)
    # continue the same solver to 1000 steps; store only the end state + observables
    refrun.launch_counts().clear()
    gl.solve.td(dt=0.1, Nt=800)
    d = {}
    d["psi1"], d["a1"], d["b1"] = state(gl)
    c = refrun.launch_counts()
    d["sweeps_psi"] = c["iterate_order_parameter_jacobi_step"] + d200["sweeps_psi"]
    d["sweeps_A"] = c["iterate_vector_potential_jacobi_step"] + d200["sweeps_A"]
    observables(gl, d, "obs_")
    np.savez_compressed(os.path.join(OUT, "cfg1_td1000.npz"), **d)
    print("cfg1_td1000 sweeps", d["sweeps_psi"], d["sweeps_A"], "vortices", d["obs_vx"].size, "E", d["obs_E"])


if __name__ == "__main__":
    main()
