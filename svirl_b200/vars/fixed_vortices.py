"""Host-side helpers for fixed (pinned) vortices.

Only the host-only part of svirl/vars/fixed_vortices.py is provided in this round
(``order_parameter_add_vortices`` and the vortex-list normalisation, used by
tests/at_precision.py:38).  The irregular vector potential / phase-lock kernels are
SURVEY.md section 8 row f1 ("next"); asking for them raises instead of silently ignoring."""
import numpy as np

import svirl_b200.config as cfg


class FixedVortices(object):

    def __init__(self, mesh, vars):
        self._vpi = None
        self._phase_lock_ns = None
        self._phase_lock_radius = cfg.phase_lock_radius
        self.mesh = mesh
        self.vars = vars
        if cfg.fixed_vortices_correction is None:
            cfg.fixed_vortices_correction = 'none'
        assert cfg.fixed_vortices_correction in ('none', 'cell centers', 'vertices')
        self.fixed_vortices_correction = cfg.fixed_vortices_correction
        vx, vy, vv = self._vortices_format(cfg.fixed_vortices)
        if vx.size > 0 or cfg.phase_lock_radius is not None:
            raise NotImplementedError("fixed vortices / phase lock are not part of this build "
                                      "(SURVEY.md section 8, row f1)")
        self.fixed_vortices_x, self.fixed_vortices_y, self.fixed_vortices_vorticity = vx, vy, vv

    @staticmethod
    def _isolated_vortex_modulus(x, y):
        r2 = x ** 2 + y ** 2
        return (1.0 - np.exp(-r2)) / (1.0 + np.exp(-r2))

    def order_parameter_add_vortices(self, vortices, phase=True, deep=False):
        """Multiply psi by a phase winding (and optionally a |psi| dip) around each vortex."""
        vx, vy, vv = self._vortices_format(vortices)
        xg, yg = self.mesh.xy_grid
        psi = self.vars._psi.get_h()
        for k in range(vx.size):
            if phase:
                psi *= np.exp(1.0j * vv[k] * np.arctan2(yg - vy[k], xg - vx[k]))
            if deep:
                psi *= np.power(self._isolated_vortex_modulus(xg - vx[k], yg - vy[k]), np.abs(vv[k]))
        self.vars._psi.need_htod_sync()
        self.vars._psi.sync()

    def _vortices_format(self, vortices):
        """Normalise (x, y[, vorticity]) given as list/tuple/dict into three equal-length arrays."""
        if vortices is None:
            vortices = [[], []]
        assert isinstance(vortices, (list, tuple, dict))
        if isinstance(vortices, dict):
            vx, vy, vv = vortices['x'], vortices['y'], vortices.get('vorticity', [])
        else:
            assert len(vortices) in [2, 3]
            vx, vy = vortices[0], vortices[1]
            vv = vortices[2] if len(vortices) == 3 else []
        vx, vy, vv = [np.array([] if v is None else v).flatten() for v in (vx, vy, vv)]
        n = max(vx.size, vy.size, vv.size)
        if vx.size > 0 and vv.size == 0:
            vv = np.array([1])
        assert vx.size in [1, n] and vy.size in [1, n] and vv.size in [1, n]
        out = []
        x = y = v = np.nan
        for i in range(n):
            x = vx[i] if i < vx.size else x
            y = vy[i] if i < vy.size else y
            v = vv[i] if i < vv.size else v
            if not (np.isnan(x) or np.isnan(y) or np.isnan(v)):
                out.append((x, y, v))
        arr = np.array(out, dtype=np.float64).reshape(-1, 3)
        return (arr[:, 0].astype(cfg.dtype), arr[:, 1].astype(cfg.dtype), arr[:, 2].astype(cfg.dtype))

    @property
    def irregular_vector_potential(self):
        return (np.zeros((cfg.Nxa, cfg.Nya), dtype=cfg.dtype), np.zeros((cfg.Nxb, cfg.Nyb), dtype=cfg.dtype))

    def irregular_vector_potential_h(self):
        return np.uintp(0)

    def _phase_lock_ns_h(self):
        return np.uintp(0)
