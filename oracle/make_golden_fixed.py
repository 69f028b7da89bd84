"""Fixed-vortex fixtures from the UNMODIFIED reference on the CPU -- TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_fixed.py        # needs /root/reference (build container only)

Writes tests/golden/td_*_fixed.npz: a TDGL run with fixed vortices and phase lock through
gl.solve.td() (both equations), then three steps through td(eqn='order_parameter') -- the path on
which the reference adds the irregular potential once and subtracts it every step -- with sweep counts,
the device copy of the irregular potential (it drifts), energy and detected vortices.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refrun  # noqa: E402
from make_golden import mt_holes, state  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def fixed_case(svirl, name, Nt, Nt2, **kw):
    refrun.launch_counts().clear()
    gl = svirl.GLSolver(**kw)
    fv = gl.params.fixed_vortices
    d = {}
    d["psi0"], d["a0"], d["b0"] = state(gl)
    ai, bi = fv.irregular_vector_potential
    d["ai0"], d["bi0"] = ai.copy(), bi.copy()
    d["fvx"], d["fvy"], d["fvv"] = fv.fixed_vortices
    d["lock_ns"] = (fv._phase_lock_ns.get_h().ravel().copy() if fv._phase_lock_ns is not None
                    else np.zeros(0, dtype=np.int32))
    if kw.get("material_tiling") is not None:
        d["mt"] = gl.mesh.material_tiling
    gl.solve.td(dt=0.1, Nt=Nt)
    c = refrun.launch_counts()
    d["sweeps_psi"] = c.get("iterate_order_parameter_jacobi_step", 0)
    d["sweeps_A"] = c.get("iterate_vector_potential_jacobi_step", 0)
    d["psi1"], d["a1"], d["b1"] = state(gl)
    d["vpi_dev1"] = np.asarray(fv._vpi.get_d_obj().get()).ravel().copy()       # packed (a_i then b_i), device copy
    ai, bi = fv.irregular_vector_potential
    d["ai1_host"], d["bi1_host"] = ai.copy(), bi.copy()
    d["obs_E"] = gl.observables.free_energy
    vx, vy, vv = gl.vortex_detector.vortices
    d["obs_vx"], d["obs_vy"], d["obs_vv"] = vx, vy, vv
    d["phase"] = fv.fixed_vortices_phase
    # observables and a few CG iterations on a COPY of the state: both see external + irregular (params.py:195-203)
    ae, be = gl.params.external_irregular_vector_potential
    d["ae"], d["be"] = ae.copy(), be.copy()
    d["obs_B"] = gl.observables.magnetic_field
    jx, jy = gl.observables.supercurrent_density
    d["obs_jsx"], d["obs_jsy"] = jx, jy
    gl.solve.td(dt=0.1, Nt=Nt2, eqn="order_parameter")
    d["psi2"], d["a2"], d["b2"] = state(gl)          # a2, b2: the host copy, which this path leaves stale (= a1, b1)
    d["vp_dev2"] = np.asarray(gl.vars._vp.get_d_obj().get()).ravel().copy()   # packed device copy (a then b)
    c = refrun.launch_counts()
    d["sweeps_psi2"] = c.get("iterate_order_parameter_jacobi_step", 0)
    d["rand_t"] = int(gl.solve._td._random_t)
    # stage three: CG from the stage-two state (host copies; the device copy of A differs, see above)
    d["psi3_in"], d["a3_in"], d["b3_in"] = state(gl)
    gl.vars.vector_potential = (d["a3_in"], d["b3_in"])          # make device = host, so the fixture is self-contained
    gl.solve.cg(n_iter=3)
    d["cg_E"] = np.array(gl.solve._cg.cg_energies, dtype=np.float64)
    d["psi3"], d["a3"], d["b3"] = state(gl)
    meta = {k: v for k, v in kw.items() if np.isscalar(v) and not callable(v) and k != "dtype"}
    meta["dtype"] = np.dtype(kw.get("dtype", np.float64)).name
    meta["Nt"], meta["Nt2"] = Nt, Nt2
    d["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "sweeps", d["sweeps_psi"], d["sweeps_A"], d["sweeps_psi2"], "vortices", d["obs_vx"].size, "E", d["obs_E"],
          "lock", d["lock_ns"].size)


def main():
    svirl = refrun.import_reference()
    base = dict(Nx=49, Ny=41, dx=0.5, dy=0.4, homogeneous_external_field=0.1, random_seed=5,
                fixed_vortices=[[8.2, 15.1], [7.3, 11.0], [1, -1]], fixed_vortices_correction="cell centers")
    fixed_case(svirl, "td_f64_k2_fixed", 12, 3, gl_parameter=2.0, normal_conductivity=10.0, phase_lock_radius=1.2,
               material_tiling=mt_holes, **base)
    fixed_case(svirl, "td_f32_kinf_fixed", 12, 3, dtype=np.float32, phase_lock_radius=1.2, **base)
    fixed_case(svirl, "td_f64_k3_fixed_nolock", 8, 2, gl_parameter=3.0, normal_conductivity=50.0, **base)


if __name__ == "__main__":
    main()
