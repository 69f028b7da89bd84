"""Solution variables: order parameter psi (nodes) and vector potential A = (a, b) (edges).
API of svirl/vars/vars.py:12-194."""
import numpy as np

from svirl_b200 import config as cfg
from svirl_b200.storage import GArray


def _edge_shapes():
    return [(cfg.Nxa, cfg.Nya), (cfg.Nxb, cfg.Nyb)]


class Vars(object):

    def __init__(self, Par, mesh):
        self.par = Par
        self.mesh = mesh
        self._psi = None
        self._vp = None
        self._tmp_node_var = None
        self._tmp_edge_var = None
        self._tmp_cell_var = None
        self._tmp_psi_real = None
        self._tmp_A_real = None
        self.solveA = bool(not np.isposinf(cfg.gl_parameter))

        if isinstance(cfg.order_parameter, str) and cfg.order_parameter == 'random':
            # the zeroing outside the material happens in the setter, i.e. BEFORE randomisation,
            # exactly as in the reference (vars.py:37-40); the first sweep zeroes inactive nodes
            self.order_parameter = 1.0
            self.randomize_order_parameter(level=cfg.random_level, seed=cfg.random_seed)
        else:
            self.order_parameter = cfg.order_parameter
        self._vp = GArray(shape=_edge_shapes(), dtype=cfg.dtype)

    # ---- order parameter
    @property
    def order_parameter(self):
        self._psi.sync()
        return self._psi.get_h().copy()

    @order_parameter.setter
    def order_parameter(self, order_parameter):
        if isinstance(order_parameter, (np.complexfloating, complex, np.floating, float, np.integer, int)):
            order_parameter = cfg.dtype_complex(order_parameter) * np.ones((cfg.Nx, cfg.Ny), cfg.dtype_complex)
        assert order_parameter.shape == (cfg.Nx, cfg.Ny)
        if self._psi is None:
            self._psi = GArray(like=np.asarray(order_parameter, dtype=cfg.dtype_complex))
        else:
            self._psi.set_h(order_parameter)
        self.set_order_parameter_to_zero_outside_material()
        self._psi.sync()

    def order_parameter_h(self):
        return self._psi.get_d_obj()

    def set_order_parameter_to_zero_outside_material(self):
        if self._psi is None or not self.mesh.have_material_tiling():
            return
        inside = self.mesh._get_material_tiling_at_nodes()
        psi = self._psi.get_h()
        psi[~inside] = 0.0
        self._psi.need_htod_sync()
        self._psi.sync()

    def randomize_order_parameter(self, level=1.0, seed=None):
        """|psi| *= 1 - level*u1 ; arg psi += level*pi*(2 u2 - 1), u uniform in [0,1): the first N
        legacy-MT draws are u1, the next N are u2 (flat, x-fastest)."""
        assert 0.0 <= level <= 1.0
        self._psi.sync()
        if seed is not None:
            np.random.seed(seed)
        modulus = 1.0 - level * np.random.rand(cfg.N)
        phase = level * 1.0j * np.pi * (2.0 * np.random.rand(cfg.N) - 1.0)
        self._psi.set_h(modulus * np.exp(phase))
        self._psi.sync()

    # ---- vector potential
    @property
    def vector_potential(self):
        if self._vp is None:
            return (np.zeros((cfg.Nxa, cfg.Nya), dtype=cfg.dtype), np.zeros((cfg.Nxb, cfg.Nyb), dtype=cfg.dtype))
        self._vp.sync()
        return self._vp.get_vec_h()

    @vector_potential.setter
    def vector_potential(self, vector_potential):
        a, b = vector_potential
        self._vp.set_vec_h(a, b)
        self._vp.sync()

    def vector_potential_h(self):
        return self._vp.get_d_obj() if self._vp is not None else np.uintp(0)

    # ---- temporaries
    def _tmp_node_var_h(self):
        if self._tmp_node_var is None:
            self._tmp_node_var = GArray(like=self._psi)
        return self._tmp_node_var.get_d_obj()

    def _tmp_edge_var_h(self):
        if self._tmp_edge_var is None:
            self._tmp_edge_var = GArray(shape=_edge_shapes(), dtype=cfg.dtype)
        return self._tmp_edge_var.get_d_obj()

    def _tmp_cell_var_h(self):
        if self._tmp_cell_var is None:
            self._tmp_cell_var = GArray(shape=(cfg.Nxc, cfg.Nyc), dtype=cfg.dtype)
        return self._tmp_cell_var.get_d_obj()

    def _tmp_psi_real_h(self):
        return self._tmp_psi_real.get_d_obj() if self._tmp_psi_real is not None else np.uintp(0)

    def _tmp_A_real_h(self):
        return self._tmp_A_real.get_d_obj() if self._tmp_A_real is not None else np.uintp(0)

    def _alloc_free_temporary_gpu_storage(self, action):
        """Kept for API parity; the library owns its reduction scratch, nothing to do."""
        assert action in ['alloc', 'free']
