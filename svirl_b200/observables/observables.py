"""Physical observables (API of svirl/observables/observables.py:10-149); the arithmetic runs
in the library (svl_free_energy, svl_magnetic_field, svl_current_density,
svl_supercurrent_density)."""
import ctypes as C

import numpy as np

import svirl_b200.config as cfg
from svirl_b200 import _lib
from svirl_b200.parallel.utils import Utils


def _h(x):
    """Device handle or NULL (the reference passes np.uintp(0) for absent arrays)."""
    return x.handle if hasattr(x, 'handle') else None


class Observables(object):

    def __init__(self, Par, mesh, vars, params):
        self.par = Par
        self.mesh = mesh
        self.vars = vars
        self.params = params

    def _eps_args(self):
        p = self.params
        return float(np.asarray(p.linear_coefficient_scalar_h()).reshape(-1)[0]), _h(p.linear_coefficient_h())

    @property
    def superfluid_density(self):
        self.vars._psi.sync()
        return Utils.abs2(self.vars._psi.get_h())

    @property
    def magnetic_field(self):
        """Induced magnetic field on cells."""
        self.vars._vp.push()
        if self.params._vpei is not None:
            self.params._vpei.push()
        _lib.call("svl_magnetic_field", self.par.ctx, _h(self.params.external_irregular_vector_potential_h()),
                  _h(self.vars.vector_potential_h()), self.vars._tmp_cell_var_h().handle)
        self.vars._tmp_cell_var.need_dtoh_sync()
        return self.vars._tmp_cell_var.get_h().copy()

    @property
    def supercurrent_density(self):
        """Superconducting current density on (horizontal, vertical) edges."""
        self.vars._psi.push()
        self.vars._vp.push()
        if self.params._vpei is not None:
            self.params._vpei.push()
        _lib.call("svl_supercurrent_density", self.par.ctx, self.vars.order_parameter_h().handle,
                  _h(self.params.external_irregular_vector_potential_h()), _h(self.vars.vector_potential_h()),
                  self.vars._tmp_edge_var_h().handle)
        self.vars._tmp_edge_var.need_dtoh_sync()
        jsx, jsy = self.vars._tmp_edge_var.get_vec_h()
        return (jsx.copy(), jsy.copy())

    @property
    def current_density(self):
        """Total current density on edges; equals the supercurrent when kappa is infinite."""
        if not self.params.solveA:
            return self.supercurrent_density
        self.vars._vp.push()
        if self.params._vpei is not None:
            self.params._vpei.push()
        _lib.call("svl_current_density", self.par.ctx, float(self.params.gl_parameter_squared_h()),
                  float(self.params.homogeneous_external_field),
                  _h(self.params.external_irregular_vector_potential_h()), _h(self.vars.vector_potential_h()),
                  self.vars._tmp_edge_var_h().handle)
        self.vars._tmp_edge_var.need_dtoh_sync()
        jx, jy = self.vars._tmp_edge_var.get_vec_h()
        return (jx.copy(), jy.copy())

    @property
    def normalcurrent_density(self):
        jx, jy = self.current_density
        jsx, jsy = self.supercurrent_density
        return (jx - jsx, jy - jsy)

    @property
    def free_energy(self):
        """Total GL free energy."""
        self.vars._psi.push()
        self.vars._vp.push()
        eps, epsf = self._eps_args()
        E = C.c_double()
        _lib.call("svl_free_energy", self.par.ctx, float(self.params.gl_parameter_squared_h()), eps, epsf,
                  float(self.params.homogeneous_external_field), self.vars.order_parameter_h().handle,
                  _h(self.params.external_irregular_vector_potential_h()), _h(self.vars.vector_potential_h()),
                  C.byref(E))
        return cfg.dtype(E.value)
