"""svirl_b200: B200-native TDGL / nonlinear-CG solver behind svirl's Python API.

    from svirl_b200 import GLSolver
    gl = GLSolver(dx=0.5, dy=0.5, Lx=64, Ly=64, gl_parameter=5.0, normal_conductivity=200.0,
                  homogeneous_external_field=0.1)
    gl.solve.td(dt=0.1, Nt=1000)
    vx, vy, vv = gl.vortex_detector.vortices

Same constructor keywords, attributes and solver calls as ``svirl.GLSolver``
(svirl/__init__.py:16-206); the device layer is libsvirl_b200.so (include/svirl_b200.h)
instead of pyCUDA.  There is no CPU fallback."""
import numpy as np

from svirl_b200 import config as cfg
from svirl_b200 import parallel as GLPar
from svirl_b200 import mesh as GLMesh
from svirl_b200 import vars as GLVars
from svirl_b200 import observables as GLObs
from svirl_b200 import solvers as GLSolvers

__all__ = ["GLSolver"]

_NUM = (np.floating, float, np.integer, int)


def _resolve_axis(N, d, L, name):
    """Two of (N, d, L) define the third: L = d*(N-1)."""
    if N is not None and L is not None and d is None:
        d = float(L) / (N - 1)
    elif N is not None and L is None and d is not None:
        L = float(d) * (N - 1)
    elif N is None and L is not None and d is not None:
        N = int(np.round(L / d) + 1)
    elif N is not None and L is not None and d is not None:
        assert np.isclose(L, d * (N - 1))
    else:
        raise ValueError('Two out of three N%s, L%s, d%s must be defined' % (name, name, name))
    assert isinstance(L, _NUM) and L > 0.0
    assert isinstance(N, (np.integer, int)) and N >= 4
    return N, d, L


class GLSolver(object):
    """2D Ginzburg-Landau solver on the gauge-invariant link-variable lattice:
    time-dependent GL (finite or infinite GL parameter), nonlinear-CG free-energy
    minimisation, user-defined material tiling, observables and a vortex detector."""

    def __init__(self,
                 Nx=None, dx=None, Lx=None, Ny=None, dy=None, Ly=None,
                 Nt=None, dt=None, T=None, NtA=None, dtA=None, TA=None,
                 material_tiling=None, order_parameter='random', random_seed=None, random_level=1.0,
                 gl_parameter=np.inf, normal_conductivity=1.0, linear_coefficient=1.0,
                 homogeneous_external_field=0.0, external_field=0.0,
                 fixed_vortices=None, fixed_vortices_correction='cell centers', phase_lock_radius=None,
                 device_id=0, dtype=np.float64,
                 stop_criterion_order_parameter=1e-6, stop_criterion_vector_potential=1e-6,
                 order_parameter_Langevin_coefficient=0.0, vector_potential_Langevin_coefficient=0.0,
                 convergence_rtol=1e-6,
                 slab=None):
        self.dtypes = (np.float32, np.float64)
        assert dtype in self.dtypes
        cfg.device_id = device_id
        cfg.dtype = dtype
        cfg.dtype_complex = {np.float32: np.complex64, np.float64: np.complex128}[dtype]

        Nx, dx, Lx = _resolve_axis(Nx, dx, Lx, 'x')
        Ny, dy, Ly = _resolve_axis(Ny, dy, Ly, 'y')
        cfg.Nx, cfg.Ny = np.int32(Nx), np.int32(Ny)
        cfg.Lx, cfg.Ly = cfg.dtype(Lx), cfg.dtype(Ly)
        cfg.dx, cfg.dy = cfg.dtype(dx), cfg.dtype(dy)
        cfg.N = cfg.Nx * cfg.Ny
        cfg.Nxa, cfg.Nya = cfg.Nx - 1, cfg.Ny            # horizontal edges
        cfg.Nxb, cfg.Nyb = cfg.Nx, cfg.Ny - 1            # vertical edges
        cfg.Na, cfg.Nb = cfg.Nxa * cfg.Nya, cfg.Nxb * cfg.Nyb
        cfg.Nab = cfg.Na + cfg.Nb
        cfg.Nxc, cfg.Nyc = cfg.Nx - 1, cfg.Ny - 1        # cells
        cfg.Nc = cfg.Nxc * cfg.Nyc
        cfg.idx, cfg.idy = 1.0 / cfg.dx, 1.0 / cfg.dy
        cfg.idx2, cfg.idy2, cfg.idxy = cfg.idx * cfg.idx, cfg.idy * cfg.idy, cfg.idx * cfg.idy
        cfg.j_dx, cfg.j_dy = 1.0j * cfg.dx, 1.0j * cfg.dy

        cfg.material_tiling = material_tiling
        cfg.order_parameter = order_parameter
        cfg.random_seed = random_seed
        cfg.random_level = random_level
        cfg.gl_parameter = gl_parameter
        cfg.linear_coefficient = linear_coefficient
        cfg.normal_conductivity = normal_conductivity
        cfg.homogeneous_external_field = homogeneous_external_field
        cfg.external_field = external_field
        cfg.fixed_vortices = fixed_vortices
        cfg.fixed_vortices_correction = fixed_vortices_correction
        cfg.phase_lock_radius = phase_lock_radius
        cfg.order_parameter_Langevin_coefficient = order_parameter_Langevin_coefficient
        cfg.vector_potential_Langevin_coefficient = vector_potential_Langevin_coefficient
        cfg.Nt, cfg.dt, cfg.T = None, None, None
        cfg.NtA, cfg.dtA, cfg.TA = None, None, None
        assert isinstance(stop_criterion_order_parameter, _NUM) and stop_criterion_order_parameter > 0.0
        assert isinstance(stop_criterion_vector_potential, _NUM) and stop_criterion_vector_potential > 0.0
        cfg.stop_criterion_order_parameter = cfg.dtype(stop_criterion_order_parameter)
        cfg.stop_criterion_vector_potential = cfg.dtype(stop_criterion_vector_potential)
        cfg.convergence_rtol = convergence_rtol
        # slab = None (whole grid on this GPU), (j0, j1), or 'auto' (rows split over the ranks of the
        # default torch.distributed group; neighbours are connected at the end of construction)
        self._auto_slab = isinstance(slab, str) and slab == 'auto'
        if self._auto_slab:
            import torch.distributed as dist
            from svirl_b200.parallel.slab import partition_rows
            slab = partition_rows(int(cfg.Ny), dist.get_world_size())[dist.get_rank()]
        cfg.slab = tuple(int(x) for x in slab) if slab is not None else None
        self.cfg = cfg

        # same construction order as the reference (svirl/__init__.py:158-175)
        self.par = GLPar.Startup()
        self.par.red = GLPar.Reduction(self.par)
        self.mesh = GLMesh.Grid()
        self.vars = GLVars.Vars(self.par, self.mesh)
        self.params = GLVars.Params(self.mesh, self.vars)
        self.observables = GLObs.Observables(self.par, self.mesh, self.vars, self.params)
        self.solve = GLSolvers.Solvers(self.par, self.mesh, self.vars, self.params, self.observables)
        self.vortex_detector = GLObs.VortexDetector(self.vars, self.params, self.solve)
        self.slab_comm = None
        if self._auto_slab:
            from svirl_b200.parallel.slab import SlabComm
            self.slab_comm = SlabComm(self)

    # ---- helpers used by the reference's tests
    def flatten_a_array(self, a):
        return np.reshape(a.T, self.cfg.Na)

    def unflatten_a_array(self, a):
        return np.reshape(a, (self.cfg.Nya, self.cfg.Nxa)).T

    def flatten_b_array(self, b):
        return np.reshape(b.T, self.cfg.Nb)

    def unflatten_b_array(self, b):
        return np.reshape(b, (self.cfg.Nyb, self.cfg.Nxb)).T

    def flatten_array(self, psi):
        return np.reshape(psi.T, self.cfg.N)

    def unflatten_array(self, psi):
        return np.reshape(psi, (self.cfg.Ny, self.cfg.Nx)).T
