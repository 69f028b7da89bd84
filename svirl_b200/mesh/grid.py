"""Grid coordinates and material tiling (API of svirl/mesh/grid.py:11-184).

The tiling is a bool array on cells, shape (Nx-1, Ny-1), True = superconductor.  Setting it
uploads the cells and rebuilds the per-node flag plane the kernels read (svl_set_material)."""
import numpy as np

import svirl_b200.config as cfg
from svirl_b200 import _lib
from svirl_b200.storage import GArray


def _axis(lo, hi, n):
    return np.linspace(lo, hi, num=n, endpoint=True, dtype=cfg.dtype)


class Grid(object):

    def __init__(self):
        self._mt = None
        self.material_tiling = cfg.material_tiling

    def have_material_tiling(self):
        return self._mt is not None

    # ---- coordinates
    @property
    def xy(self):
        """Node coordinates."""
        return _axis(0.0, cfg.Lx, cfg.Nx), _axis(0.0, cfg.Ly, cfg.Ny)

    @property
    def xy_grid(self):
        return np.meshgrid(*self.xy, indexing='ij')

    @property
    def xy_a(self):
        """Mid-points of the horizontal (a) edges."""
        return _axis(0.5 * cfg.dx, cfg.Lx - 0.5 * cfg.dx, cfg.Nxa), _axis(0.0, cfg.Ly, cfg.Nya)

    @property
    def xy_a_grid(self):
        return np.meshgrid(*self.xy_a, indexing='ij')

    @property
    def xy_b(self):
        """Mid-points of the vertical (b) edges."""
        return _axis(0.0, cfg.Lx, cfg.Nxb), _axis(0.5 * cfg.dy, cfg.Ly - 0.5 * cfg.dy, cfg.Nyb)

    @property
    def xy_b_grid(self):
        return np.meshgrid(*self.xy_b, indexing='ij')

    @property
    def xy_c(self):
        """Cell centres."""
        return (_axis(0.5 * cfg.dx, cfg.Lx - 0.5 * cfg.dx, cfg.Nxc),
                _axis(0.5 * cfg.dy, cfg.Ly - 0.5 * cfg.dy, cfg.Nyc))

    @property
    def xy_c_grid(self):
        return np.meshgrid(*self.xy_c, indexing='ij')

    # ---- material tiling
    @property
    def material_tiling(self):
        if self._mt is not None:
            return self._mt.get_h().copy()
        return np.full((cfg.Nxc, cfg.Nyc), True, dtype=bool)

    @material_tiling.setter
    def material_tiling(self, material_tiling):
        from svirl_b200.parallel import startup
        mt = material_tiling(*self.xy_c_grid) if callable(material_tiling) else material_tiling
        if self._mt is not None:
            self._mt.free()
            self._mt = None
        if mt is not None:
            assert mt.shape == (cfg.Nxc, cfg.Nyc)
            self._mt = GArray(like=np.asarray(mt).astype(bool))
        par = startup.active()
        _lib.call("svl_set_material", par.ctx, self._mt.get_d_obj().handle if self._mt is not None else None)

    def material_tiling_h(self):
        return self._mt.get_d_obj() if self._mt is not None else np.uintp(0)

    def _get_material_tiling_at_nodes(self):
        """True where at least one of the four cells around the node is material."""
        mt = self._mt.get_h()
        P = np.zeros((cfg.Nx + 1, cfg.Ny + 1), dtype=bool)
        P[1:-1, 1:-1] = mt
        return P[:-1, :-1] | P[:-1, 1:] | P[1:, :-1] | P[1:, 1:]

    def interpolate_ab_array_to_c_array(self, a, b):
        return 0.5 * (a[:, :-1] + a[:, 1:]), 0.5 * (b[:-1, :] + b[1:, :])

    def interpolate_ab_array_to_c_array_abs(self, a, b):
        cx, cy = self.interpolate_ab_array_to_c_array(a, b)
        return np.sqrt(np.square(cx) + np.square(cy))
