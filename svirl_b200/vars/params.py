"""Material and field parameters (API of svirl/vars/params.py:10-327)."""
import numpy as np

import svirl_b200.config as cfg
from svirl_b200.storage import GArray
from .fixed_vortices import FixedVortices

_NUM = (np.floating, float, np.integer, int)


class Params(object):

    def __init__(self, mesh, vars):
        self.mesh = mesh
        self.vars = vars
        self.fixed_vortices = FixedVortices(self.mesh, self.vars)
        self.solveA = False
        self.linear_coefficient = cfg.linear_coefficient
        self.gl_parameter = cfg.gl_parameter
        self.normal_conductivity = cfg.normal_conductivity
        self._H = cfg.dtype(0.0)
        self.homogeneous_external_field_reset = cfg.homogeneous_external_field
        self.ae, self.be = None, None
        self._vpei = None          # external + irregular potential; a zero array by default (quirk Q4)
        self.external_field = cfg.external_field
        self.order_parameter_Langevin_coefficient = cfg.order_parameter_Langevin_coefficient
        self.vector_potential_Langevin_coefficient = cfg.vector_potential_Langevin_coefficient

    # ---- linear coefficient epsilon: scalar or (Nx, Ny) field
    @property
    def linear_coefficient(self):
        if self._epsilon.size == 1:
            return np.full((cfg.Nx, cfg.Ny), self._epsilon.get_h(), dtype=cfg.dtype)
        return self._epsilon.get_h()

    @linear_coefficient.setter
    def linear_coefficient(self, linear_coefficient):
        lc = linear_coefficient(*self.mesh.xy_grid) if callable(linear_coefficient) else linear_coefficient
        if np.isscalar(lc):
            lc = lc * np.ones(1)
        else:
            assert lc.shape == (cfg.Nx, cfg.Ny)
        self._epsilon = GArray(like=lc.astype(cfg.dtype))

    def linear_coefficient_h(self):
        return self._epsilon.get_d_obj() if self._epsilon.size != 1 else np.uintp(0)

    def linear_coefficient_scalar_h(self):
        # 0.0 when epsilon is a field (params.py:79-83) -- this is what makes quirk Q11 visible
        return self._epsilon.get_h() if self._epsilon.size == 1 else cfg.dtype(0.0)

    # ---- GL parameter kappa
    @property
    def gl_parameter(self):
        return self._kappa

    @gl_parameter.setter
    def gl_parameter(self, gl_parameter):
        if gl_parameter is None or np.isnan(gl_parameter) or np.isinf(gl_parameter):
            gl_parameter = np.inf
        assert isinstance(gl_parameter, _NUM) and (np.isposinf(gl_parameter) or gl_parameter > 0.0)
        self._kappa = cfg.dtype(gl_parameter)
        self.solveA = bool(not np.isposinf(self._kappa))

    def gl_parameter_squared_h(self):
        return cfg.dtype(self.gl_parameter ** 2) if self.solveA else cfg.dtype(-1.0)

    # ---- conductivity
    @property
    def normal_conductivity(self):
        return self._sigma

    @normal_conductivity.setter
    def normal_conductivity(self, normal_conductivity):
        assert isinstance(normal_conductivity, _NUM) and normal_conductivity > 0.0
        self._sigma = cfg.dtype(normal_conductivity)
        self._rho = cfg.dtype(1.0 / normal_conductivity)

    # ---- homogeneous field H
    @property
    def homogeneous_external_field(self):
        return self._H

    @homogeneous_external_field.setter
    def homogeneous_external_field(self, homogeneous_external_field):
        self._H = cfg.dtype(homogeneous_external_field)

    def _update_vector_potential(self, homogeneous_external_field, reset):
        """Symmetric gauge on edge mid-points: a -= (y - Ly/2) dH / 2, b += (x - Lx/2) dH / 2."""
        assert isinstance(homogeneous_external_field, _NUM)
        vp = self.vars._vp
        if reset:
            self._H = cfg.dtype(homogeneous_external_field)
            a, b = vp.get_vec_h()
            a.fill(0.0)
            b.fill(0.0)
            vp.need_htod_sync()
            vp.sync()
            delta_H = self._H
        else:
            delta_H = - self._H
            self._H = cfg.dtype(homogeneous_external_field)
            delta_H += self._H
            vp.sync()
        g = 0.5
        _, yg = self.mesh.xy_a_grid
        xg, _ = self.mesh.xy_b_grid
        a, b = vp.get_vec_h()
        a -= g * (yg - 0.5 * cfg.Ly) * delta_H
        b += (1.0 - g) * (xg - 0.5 * cfg.Lx) * delta_H
        vp.need_htod_sync()
        vp.sync()

    def _homogeneous_external_field_delta(self, homogeneous_external_field):
        self._update_vector_potential(homogeneous_external_field, reset=False)

    homogeneous_external_field_delta = property(
        fset=_homogeneous_external_field_delta,
        doc="Set H and add dA with curl(dA) = H - H_old to the vector potential.")

    def _homogeneous_external_field_reset(self, homogeneous_external_field):
        self._update_vector_potential(homogeneous_external_field, reset=True)

    homogeneous_external_field_reset = property(
        fset=_homogeneous_external_field_reset,
        doc="Set H and reset the vector potential to the symmetric gauge with curl(A) = H.")

    # ---- external (non-homogeneous) + irregular potential
    def _update_gvpei(self):
        assert (self.ae is None) == (self.be is None)
        ai = bi = None
        if self.fixed_vortices is not None and self.fixed_vortices._vpi is not None:
            ai, bi = self.fixed_vortices._vpi.get_vec_h()
        if self.ae is not None:
            vpei = (self.ae + ai, self.be + bi) if ai is not None else (self.ae, self.be)
        else:
            vpei = (ai, bi) if ai is not None else None
        if self._vpei is not None:
            self._vpei.free()
            self._vpei = None
        self._vpei_is_zero = True
        if vpei is not None:
            self._vpei = GArray(shape=[vpei[0].shape, vpei[1].shape], dtype=cfg.dtype)
            self._vpei.set_vec_h(vpei[0], vpei[1])
            self._vpei.sync()
            self._vpei_is_zero = not (np.any(vpei[0]) or np.any(vpei[1]))

    @property
    def external_vector_potential(self):
        assert (self.ae is None) == (self.be is None)
        return (self.ae, self.be) if self.ae is not None else None

    @external_vector_potential.setter
    def external_vector_potential(self, external_vector_potential):
        Ax = Ay = None
        if external_vector_potential is not None:
            Ax, Ay = external_vector_potential
            assert (Ax is None) == (Ay is None)
        if Ax is not None:
            assert Ax.shape == (cfg.Nxa, cfg.Nya) and Ay.shape == (cfg.Nxb, cfg.Nyb)
        self.ae, self.be = Ax, Ay
        self._update_gvpei()

    @property
    def external_irregular_vector_potential(self):
        return self._vpei.get_vec_h() if self._vpei is not None else None

    def external_irregular_vector_potential_h(self):
        return self._vpei.get_d_obj() if self._vpei is not None else np.uintp(0)

    def _external_irregular_for_kernels(self):
        """The buffer the CG kernels add to the regular potential, or NULL when it is identically
        zero (the reference keeps a zero array by default, quirk Q4; adding 0.0 changes nothing,
        and skipping it saves two plane reads per kernel)."""
        if self._vpei is None or getattr(self, '_vpei_is_zero', False):
            return np.uintp(0)
        return self._vpei.get_d_obj()

    @property
    def external_field(self):
        A = self.external_vector_potential
        if A is None:
            return None
        Ax, Ay = A
        return - np.diff(Ax, axis=1) * cfg.idy + np.diff(Ay, axis=0) * cfg.idx

    @external_field.setter
    def external_field(self, external_field):
        if external_field is None:
            self.external_vector_potential = None
            return
        g = 0.5
        _, yg = self.mesh.xy_a_grid
        xg, _ = self.mesh.xy_b_grid
        Ax = - g * (yg - 0.5 * cfg.Ly) * external_field
        Ay = (1.0 - g) * (xg - 0.5 * cfg.Lx) * external_field
        self.external_vector_potential = (Ax, Ay)

    # ---- Langevin coefficients
    @property
    def order_parameter_Langevin_coefficient(self):
        return self._psi_langevin_c

    @order_parameter_Langevin_coefficient.setter
    def order_parameter_Langevin_coefficient(self, value):
        assert isinstance(value, _NUM)
        self._psi_langevin_c = cfg.dtype(value)

    @property
    def vector_potential_Langevin_coefficient(self):
        return self._ab_langevin_c

    @vector_potential_Langevin_coefficient.setter
    def vector_potential_Langevin_coefficient(self, value):
        assert isinstance(value, _NUM)
        self._ab_langevin_c = cfg.dtype(value)
