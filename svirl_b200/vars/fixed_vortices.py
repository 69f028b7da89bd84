"""Fixed (pinned) vortices: irregular vector potential, phase lock, release.

API of svirl/vars/fixed_vortices.py:9-246.  A fixed vortex of vorticity v at (x0, y0) is a
singular gauge field: the "irregular" potential A_i = v * grad(atan2(y - y0, x - x0)) taken as finite
differences of the angle on the edges, which the solvers add to the regular potential.  The
reference's quirks are kept: the phase-lock list holds the i-indices of the
locked nodes only (fixed_vortices.py:236), i.e. flat node numbers in the first grid row."""
import numpy as np

import svirl_b200.config as cfg
from svirl_b200.storage import GArray


class FixedVortices(object):

    def __init__(self, mesh, vars):
        self._vpi = None               # irregular potential (a_i, b_i) on the edges
        self._phase_lock_ns = None     # int32 list of locked flat node indices
        self._phase_lock_radius = cfg.phase_lock_radius
        self.mesh = mesh
        self.vars = vars
        if cfg.fixed_vortices_correction is None:
            cfg.fixed_vortices_correction = 'none'
        assert cfg.fixed_vortices_correction in ('none', 'cell centers', 'vertices')
        self.fixed_vortices_correction = cfg.fixed_vortices_correction
        self.fixed_vortices = cfg.fixed_vortices

    # ---- vortex list
    @property
    def fixed_vortices(self):
        return (self.fixed_vortices_x.copy(), self.fixed_vortices_y.copy(), self.fixed_vortices_vorticity.copy())

    @fixed_vortices.setter
    def fixed_vortices(self, vortices):
        self.fixed_vortices_x, self.fixed_vortices_y, self.fixed_vortices_vorticity = self._vortices_format(vortices)
        self._snap_to_grid()
        if self.fixed_vortices_x.size == 0:
            if self._vpi is not None:
                self._vpi.free()
                self._vpi = None
        else:
            if self._vpi is None:
                self._vpi = GArray(shape=[(cfg.Nxa, cfg.Nya), (cfg.Nxb, cfg.Nyb)], dtype=cfg.dtype)
            ai, bi = self._vpi.get_vec_h()
            ai[...] = 0.0
            bi[...] = 0.0
            xg, yg = self.mesh.xy_grid
            for x0, y0, v in zip(self.fixed_vortices_x, self.fixed_vortices_y, self.fixed_vortices_vorticity):
                theta = np.arctan2(yg - y0, xg - x0)
                theta -= theta[0, 0]
                ai += v * cfg.idx * (theta[1:, :] - theta[:-1, :])       # d(theta)/dx on the a-edges
                bi += v * cfg.idy * (theta[:, 1:] - theta[:, :-1])       # d(theta)/dy on the b-edges
            self._vpi.need_htod_sync()
            self._vpi.sync()
        self._rebuild_phase_lock()
        # like the reference, this does NOT refresh Params' external+irregular sum (params.py:195-203 is only
        # run by the external-potential setters), so CG / observables keep the sum made at construction

    def _snap_to_grid(self):
        """Integer vorticity; positions moved to cell centres or grid vertices if asked for
        (fixed_vortices.py:114-127)."""
        self.fixed_vortices_vorticity = np.round(self.fixed_vortices_vorticity)
        if self.fixed_vortices_correction == 'cell centers':
            self.fixed_vortices_x = cfg.dx * (np.round(self.fixed_vortices_x / cfg.dx + 0.5) - 0.5)
            self.fixed_vortices_y = cfg.dy * (np.round(self.fixed_vortices_y / cfg.dy + 0.5) - 0.5)
        elif self.fixed_vortices_correction == 'vertices':
            self.fixed_vortices_x = cfg.dx * np.round(self.fixed_vortices_x / cfg.dx)
            self.fixed_vortices_y = cfg.dy * np.round(self.fixed_vortices_y / cfg.dy)

    # ---- phase lock
    @property
    def phase_lock_radius(self):
        return self._phase_lock_radius

    @phase_lock_radius.setter
    def phase_lock_radius(self, radius):
        assert radius is None or (isinstance(radius, (np.floating, float, np.integer, int)) and radius > 0.0)
        self._phase_lock_radius = radius
        self._rebuild_phase_lock()

    def _rebuild_phase_lock(self):
        if self._phase_lock_ns is not None:
            self._phase_lock_ns.free()
            self._phase_lock_ns = None
        if self._phase_lock_radius is None:
            return
        xg, yg = self.mesh.xy_grid
        locked = np.zeros((cfg.Nx, cfg.Ny), dtype=bool)
        for x0, y0 in zip(self.fixed_vortices_x, self.fixed_vortices_y):
            locked |= (xg - x0) ** 2 + (yg - y0) ** 2 <= self._phase_lock_radius ** 2
        # the reference keeps the FIRST index array of np.where only, i.e. the i-index of every locked
        # node (fixed_vortices.py:236); used as flat node numbers these lie in grid row j = 0
        ns = np.nonzero(locked)[0].astype(np.int32)
        if ns.size > 0:
            self._phase_lock_ns = GArray(like=ns)

    def _phase_lock_ns_h(self):
        return self._phase_lock_ns.get_d_obj() if self._phase_lock_ns is not None else np.uintp(0)

    # ---- irregular potential and its phase
    @property
    def irregular_vector_potential(self):
        if self._vpi is None:
            return (np.zeros((cfg.Nxa, cfg.Nya), dtype=cfg.dtype), np.zeros((cfg.Nxb, cfg.Nyb), dtype=cfg.dtype))
        self._vpi.sync()
        return self._vpi.get_vec_h()

    def irregular_vector_potential_h(self):
        return self._vpi.get_d_obj() if self._vpi is not None else np.uintp(0)

    @staticmethod
    def _ab_phase(a, b):
        """Phase whose lattice gradient is (a, b), with phase[0, 0] = 0: integrate b along the first
        grid column, then a along every row (fixed_vortices.py:130-137)."""
        along_y = cfg.dy * np.concatenate([np.zeros((cfg.Nxb, 1), dtype=cfg.dtype), b], axis=1).cumsum(axis=1)
        along_x = cfg.dx * np.concatenate([np.zeros((1, cfg.Nya), dtype=cfg.dtype), a], axis=0).cumsum(axis=0)
        return along_y[0:1, :] + along_x

    @property
    def fixed_vortices_phase(self):
        ai, bi = self.irregular_vector_potential
        return self._ab_phase(ai, bi)

    def fixed_vortices_release(self):
        """Turn the fixed vortices into natural ones: move their phase winding into psi."""
        self.vars._psi.sync()
        psi = self.vars._psi.get_h()
        psi *= np.exp(-1.0j * self.fixed_vortices_phase)
        self.vars._psi.need_htod_sync()
        self.vars._psi.sync()
        self.fixed_vortices = None
        self.phase_lock_radius = None

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
