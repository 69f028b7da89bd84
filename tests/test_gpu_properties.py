"""GPU: the reference's own acceptance properties (svirl tests/at_*.py), restated as seeded pytest cases
against this package's public and private API -- the same properties and tolerances, not the same code:

  * at_cg_jacobians.py : dG/dpsi and dG/dA equal finite differences of the free energy (h = 3e-9,
                         atol 1e-5, rtol 1e-3), random size / spacing / kappa / fields / material tilings;
  * at_cg_coef.py, at_cg_coef_psi.py : the line-search polynomial reproduces the free energy along the
                         search direction, and truncating it in alpha_A at order 0 / 2 / 4 only gets better;
  * at_precision.py    : a long fp32 run lands within 10 % of the fp64 energy;
  * at_destructor.py   : solvers of random shape can be built and destroyed repeatedly.
"""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

TILINGS = ["full", "empty", "q1", "q2", "q3", "q4", "none", "random10", "random30", "random50", "random90",
           "random200", "random400", "random800"]


def random_tiling(gl, rs, kind):
    """The material tilings the reference's tests draw from (tests/common.py:14-62): full, empty, one
    quadrant removed, or p * Nc randomly chosen cells removed (with repetition)."""
    Nxc, Nyc = int(gl.cfg.Nxc), int(gl.cfg.Nyc)
    mt = np.ones((Nxc, Nyc), dtype=bool)
    if kind == "empty":
        mt[:] = False
    elif kind.startswith("random"):
        for _ in range(int(0.01 * int(kind[6:]) * Nxc * Nyc)):
            mt[rs.randint(Nxc), rs.randint(Nyc)] = False
    elif kind.startswith("q"):
        h = Nxc // 2
        sl = {"q1": (slice(None, h), slice(None, h)), "q2": (slice(h, None), slice(None, h)),
              "q3": (slice(None, h), slice(h, None)), "q4": (slice(h, None), slice(h, None))}[kind]
        mt[sl] = False
    elif kind == "none":
        mt = None
    gl.mesh.material_tiling = mt
    gl.vars.set_order_parameter_to_zero_outside_material()


def random_state(gl, rs, kind):
    gl.params.gl_parameter = 1.0 + 3.0 * rs.rand()
    gl.vars.order_parameter = 1.0
    gl.vars.randomize_order_parameter(level=0.5, seed=int(rs.randint(1 << 30)))
    gl.params.homogeneous_external_field_reset = 0.01 + 0.1 * rs.rand()      # curl a = H
    gl.params.external_field = 0.01 + 0.1 * rs.rand()
    random_tiling(gl, rs, kind)


def small_solver(rs):
    from svirl_b200 import GLSolver
    return GLSolver(Nx=8 + rs.randint(4), Ny=8 + rs.randint(4), dx=0.5 - 0.1 * rs.rand(), dy=0.5 - 0.1 * rs.rand(),
                    gl_parameter=1.0)


@pytest.mark.parametrize("seed", range(6))
def test_jacobians_equal_finite_differences(seed):
    rs = np.random.RandomState(100 + seed)
    gl = small_solver(rs)
    gl.solve._init_cg()
    for kind in (TILINGS[(2 * seed) % len(TILINGS)], TILINGS[(2 * seed + 7) % len(TILINGS)]):
        random_state(gl, rs, kind)
        h = 3e-9
        E0 = gl.observables.free_energy
        psi = gl.vars.order_parameter
        fd_psi = np.zeros_like(psi)
        for i in range(psi.shape[0]):
            for j in range(psi.shape[1]):
                for unit in (1.0, 1.0j):
                    p = psi.copy()
                    p[i, j] += unit * h
                    gl.vars.order_parameter = p
                    fd_psi[i, j] += unit * (gl.observables.free_energy - E0) / h
        gl.vars.order_parameter = psi
        # the setter zeroes psi outside the material, where the Jacobian is zero as well
        jac_psi = gl.unflatten_array(gl.solve._cg._free_energy_jacobian_psi.get())
        assert np.allclose(jac_psi, fd_psi, atol=1e-5, rtol=1e-3), kind
        a, b = gl.vars.vector_potential
        fd_a, fd_b = np.zeros_like(a), np.zeros_like(b)
        for arr, fd, which in ((a, fd_a, 0), (b, fd_b, 1)):
            for i in range(arr.shape[0]):
                for j in range(arr.shape[1]):
                    arr[i, j] += h
                    gl.vars.vector_potential = (a, b)
                    fd[i, j] = (gl.observables.free_energy - E0) / h
                    arr[i, j] -= h
        gl.vars.vector_potential = (a, b)
        jA = gl.solve._cg._free_energy_jacobian_A.get()
        ja, jb = gl.unflatten_a_array(jA[:gl.cfg.Na]), gl.unflatten_b_array(jA[gl.cfg.Na:])
        assert np.allclose(ja, fd_a, atol=1e-5, rtol=1e-3) and np.allclose(jb, fd_b, atol=1e-5, rtol=1e-3), kind
    gl.par.close()


@pytest.mark.parametrize("seed", range(4))
def test_line_search_polynomial_reproduces_energy(seed):
    from svirl_b200.storage import GArray
    rs = np.random.RandomState(200 + seed)
    gl = small_solver(rs)
    gl.solve._init_cg()
    cgs = gl.solve._cg
    psi0 = gl.vars.order_parameter
    ab0 = gl.vars.vector_potential
    dpsi = GArray(like=psi0)
    dab = GArray(shape=[ab0[0].shape, ab0[1].shape], dtype=gl.cfg.dtype)
    P = np.polynomial.polynomial
    for r in (0.0001, 0.001, 0.01, 0.1, 0.3, 1.0):
        # psi = a_psi psi0, dpsi = b_psi psi0 with a_psi + b_psi alpha_psi = 1 (same for A with a small step b_A):
        # the polynomial at (alpha_psi, alpha_A) must give back the energy of (psi0, A0)
        for j, (a_psi, b_psi, al_psi, a_A, b_A, al_A) in enumerate((
                (1.0, 0.0, 0.0, 1.0 - r, r, 1.0), (0.5, 0.5, 1.0, 1.0 - r, r, 1.0),
                (0.6976, 0.72, 0.42, 1.0 - r, r, 1.0), (0.7923, 0.31, 0.67, 1.0 - 0.6 * r, 0.6 * r, 1.0))):
            random_state(gl, rs, TILINGS[rs.randint(len(TILINGS))])
            psi0 = gl.vars.order_parameter
            ab0 = [x.copy() for x in gl.vars.vector_potential]     # the getter hands out views of the host mirror
            E0 = gl.observables.free_energy
            gl.vars.order_parameter = psi0 * a_psi
            dpsi.set_h(psi0 * b_psi)
            dpsi.sync()
            gl.vars.vector_potential = (ab0[0] * a_A, ab0[1] * a_A)
            dab.set_vec_h(ab0[0] * b_A, ab0[1] * b_A)
            dab.sync()
            c = np.array(cgs._free_energy_conjgrad_coef(dpsi.get_d_obj(), dab.get_d_obj()), dtype=np.float64)
            E1 = gl.observables.free_energy
            trunc = lambda order: P.polyval2d(al_psi, al_A, c * (np.arange(5)[None, :] <= order))
            assert np.isclose(trunc(4), E0), (r, j)
            if j == 0:
                assert np.isclose(trunc(0), E1), r                 # alpha_psi = 0, no A step: the energy of the current state
            err = np.abs(np.array([trunc(0), trunc(2), trunc(4)]) - E0)
            err[err < 1e-9] = 0.0
            assert np.all(np.diff(err) < 1e-14), (r, j, err)      # higher order in alpha_A is never worse
            gl.vars.vector_potential = ab0
            gl.vars.order_parameter = psi0
    gl.par.close()


@pytest.mark.parametrize("seed", range(3))
def test_psi_line_search_quartic_reproduces_energy(seed):
    """at_cg_coef_psi.py: infinite kappa, the 5 coefficients of G(psi + alpha dpsi) give back the energy."""
    from svirl_b200 import GLSolver
    from svirl_b200.storage import GArray
    rs = np.random.RandomState(300 + seed)
    gl = GLSolver(Nx=8 + rs.randint(4), Ny=8 + rs.randint(4), dx=0.5 - 0.1 * rs.rand(), dy=0.5 - 0.1 * rs.rand())
    gl.solve._init_cg()
    dpsi = GArray(like=gl.vars.order_parameter)
    for a_psi, b_psi, al_psi in ((1.0, 0.0, 0.0), (0.5, 0.5, 1.0), (0.6976, 0.72, 0.42), (0.7923, 0.31, 0.67)):
        gl.vars.order_parameter = 1.0
        gl.vars.randomize_order_parameter(level=0.5, seed=int(rs.randint(1 << 30)))
        gl.params.homogeneous_external_field_reset = 0.01 + 0.1 * rs.rand()
        gl.params.external_field = 0.01 + 0.1 * rs.rand()
        random_tiling(gl, rs, TILINGS[rs.randint(len(TILINGS))])
        psi0 = gl.vars.order_parameter
        E0 = gl.observables.free_energy
        gl.vars.order_parameter = psi0 * a_psi
        dpsi.set_h(psi0 * b_psi)
        dpsi.sync()
        c = np.array(gl.solve._cg._free_energy_conjgrad_coef_psi(dpsi.get_d_obj()), dtype=np.float64)
        c = c[0] if c.ndim == 2 else c
        assert np.isclose(np.polynomial.polynomial.polyval(al_psi, c), E0), (a_psi, b_psi)
    gl.par.close()


def test_fp32_run_lands_within_ten_percent_of_fp64():
    """at_precision.py at a quarter of its length (1500 steps): one vortex seeded in a 100 x 100 sample with a
    10 x 10 hole, kappa 3.6432, sigma 400, H 0.1."""
    from svirl_b200 import GLSolver
    E = {}
    for dtype in (np.float32, np.float64):
        gl = GLSolver(Lx=100, Ly=100, dx=1.0, dy=1.0, order_parameter=1.0, gl_parameter=3.6432, normal_conductivity=400.0,
                      homogeneous_external_field=0.1, dtype=dtype, convergence_rtol=1e-12)
        gl.params.fixed_vortices.order_parameter_add_vortices([50, 50], phase=True, deep=True)
        Lx, Ly = float(gl.cfg.Lx), float(gl.cfg.Ly)
        gl.mesh.material_tiling = lambda x, y: ~((np.abs(x - Lx / 2) < 5.0) & (np.abs(y - Ly / 2) < 5.0))
        gl.vars.set_order_parameter_to_zero_outside_material()
        gl.solve.td(Nt=1500, dt=0.1)
        E[dtype] = gl.observables.free_energy
        gl.par.close()
    assert np.isclose(E[np.float32], E[np.float64], rtol=1e-1), E


def test_construct_and_destroy_random_solvers():
    from svirl_b200 import GLSolver
    rs = np.random.RandomState(7)
    for _ in range(10):
        finite = rs.rand() > 0.5
        gl = GLSolver(Nx=int(rs.randint(4, 1024)), Ny=int(rs.randint(4, 1024)), dx=0.2 + 0.2 * rs.rand(),
                      dy=0.2 + 0.2 * rs.rand(), gl_parameter=1.0 if finite else np.inf)
        gl.vars.order_parameter = 1.0
        gl.vars.randomize_order_parameter(level=0.5)
        if finite:
            gl.params.gl_parameter = 1.0 + 3.0 * rs.rand()
            gl.params.external_field = 0.01 + 0.1 * rs.rand()
        gl.params.homogeneous_external_field = 0.01 + 0.1 * rs.rand()
        random_tiling(gl, rs, TILINGS[rs.randint(len(TILINGS))])
        E = gl.observables.free_energy
        assert np.isfinite(E)
        gl.par.close()
        del gl
