"""Vortex detector (API of svirl/observables/vortex_detector.py:8-156).

The reference loops over all Nc cells in the Python interpreter.  Here the GPU emits a
SUPERSET of candidate cells (svl_vortex_candidates, winding number in double with loosened
thresholds) and only those cells go through the host arithmetic below, which follows the
reference's scalar expressions term by term (same numpy scalar types, same order), so the
result is identical for identical fields; order is ascending cell index n = i + Nxc*j."""
import ctypes as C

import numpy as np

import svirl_b200.config as cfg
from svirl_b200 import _lib


def _cell_zero(x1, y1, f1, x2, y2, f2):
    """Linear-interpolation zero of f on the segment (x1,y1)-(x2,y2)."""
    return (f2 * x1 - x2 * f1) / (f2 - f1), (f2 * y1 - y2 * f1) / (f2 - f1)


def _border_zeros(c1, c2, c3, c4):
    """Zeros of the linearised f on the four cell edges, corners given as (x, y, f)."""
    out = []
    for p, q in ((c2, c1), (c3, c2), (c4, c3), (c1, c4)):
        if p[2] * q[2] < -1e-10:
            out.append(_cell_zero(p[0], p[1], p[2], q[0], q[1], q[2]))
    return out


def _lines_cross(p1, p2, q1, q2):
    """Intersection of line p1-p2 with line q1-q2 and the angle between them (folded to [0, pi/2])."""
    ux, uy = p1[0] - p2[0], p1[1] - p2[1]
    vx, vy = q1[0] - q2[0], q1[1] - q2[1]
    D = ux * vy - uy * vx
    ph = np.mod(np.abs(np.arctan2(D, uy * vy - ux * vx)), 0.5 * np.pi)
    if np.abs(ph) > 1e-10:
        cp, cq = p1[0] * p2[1] - p1[1] * p2[0], q1[0] * q2[1] - q1[1] * q2[0]
        return (cp * vx - ux * cq) / D, (cp * vy - uy * cq) / D, ph
    return np.nan, np.nan, ph


# above this many candidate cells the per-candidate arithmetic runs on whole arrays (triangulate.py)
VECTOR_THRESHOLD = 4096


class VortexDetector(object):

    def __init__(self, _vars, params, solver):
        self.vars = _vars
        self.params = params
        self.fixed_vortices = self.params.fixed_vortices
        self.solver = solver

    def unflatten_c(self, n):
        return n % cfg.Nxc, n // cfg.Nxc

    def flatten_c(self, i, j):
        return i + cfg.Nxc * j

    def _candidates(self):
        par = self.vars.par
        cap = 1 << 16
        while True:
            cells = np.empty(cap, dtype=np.int64)
            vals = np.empty(cap, dtype=np.float64)
            cnt = C.c_size_t()
            _lib.call("svl_vortex_candidates", par.ctx, float(self.params.homogeneous_external_field),
                      self.vars.order_parameter_h().handle, self.vars.vector_potential_h().handle,
                      cells.ctypes.data_as(C.POINTER(C.c_int64)), vals.ctypes.data_as(C.POINTER(C.c_double)),
                      cap, C.byref(cnt))
            if cnt.value <= cap:
                return np.sort(cells[:cnt.value])
            cap = int(cnt.value)

    @property
    def vortices(self):
        """(x, y, vorticity) of every detected vortex, sub-cell precision."""
        if self.solver.vortices_detected is not True:
            self.vars._psi.sync()
            self.vars._vp.sync()
            a, b = self.vars._vp.get_vec_h()
            a_ai, b_bi = a, b
            if self.fixed_vortices._vpi is not None:      # triangulation uses regular + irregular (vortex_detector.py:43-47)
                ai, bi = self.fixed_vortices.irregular_vector_potential
                a_ai, b_bi = a + ai, b + bi
            psi = self.vars.order_parameter
            dx, dy, pi = cfg.dx, cfg.dy, np.pi
            H = self.params.homogeneous_external_field
            found = []
            cand = self._candidates()
            if cand.size > VECTOR_THRESHOLD:
                # many vortices: the same arithmetic on whole arrays (bit-identical, tests/test_triangulate_host.py)
                from .triangulate import triangulate
                tx, ty, tv = triangulate(cand, psi, a, b, a_ai, b_bi, H, dx, dy, cfg.Nxc, 0, cfg.dtype)
                self._detected_vortices = np.stack([tx, ty, tv], axis=1).astype(cfg.dtype).reshape(-1, 3)
                self.solver.vortices_detected = True
                d = self._detected_vortices
                return d[:, 0], d[:, 1], d[:, 2]
            for n in cand:
                i, j = self.unflatten_c(int(n))   # np.int32 like the reference: dx*i promotes to float64
                ip, jp = i + 1, j + 1
                # theta = np.angle(psi) at the four corners only (the reference takes it over the whole grid,
                # vortex_detector.py:50; elementwise, so the values are the same)
                t_00, t_p0, t_pp, t_0p = (np.angle(psi[i, j]), np.angle(psi[ip, j]), np.angle(psi[ip, jp]),
                                          np.angle(psi[i, jp]))
                v = - (0.5 / pi) * (
                    np.mod(t_p0 - t_00 - dx * a[i, j] + pi, 2.0 * pi)
                    + np.mod(t_pp - t_p0 - dy * b[ip, j] + pi, 2.0 * pi)
                    + np.mod(t_0p - t_pp + dx * a[i, jp] + pi, 2.0 * pi)
                    + np.mod(t_00 - t_0p + dy * b[i, j] + pi, 2.0 * pi)
                    - 4.0 * pi
                    + dx * dy * H)
                if not (np.abs(v) > 0.5 and np.abs(v - np.round(v)) < 0.1):
                    continue
                x, y = dx * i, dy * j
                ia00, ia0p = dx * a_ai[i, j], dx * a_ai[i, jp]
                ib00, ibp0 = dy * b_bi[i, j], dy * b_bi[ip, j]
                # gauge-transport the other three corners to the corner (i, j)
                q00 = psi[i, j]
                qp0 = psi[ip, j] * np.exp(-1j * (0.75 * ia00 + 0.25 * (ib00 + ia0p - ibp0)))
                qpp = psi[ip, jp] * np.exp(-1j * (0.5 * (ia00 + ibp0) + 0.5 * (ib00 + ia0p)))
                q0p = psi[i, jp] * np.exp(-1j * (0.75 * ib00 + 0.25 * (ia00 + ibp0 - ia0p)))
                xs, ys = (x, x + dx, x + dx, x), (y, y, y + dy, y + dy)
                qs = (q00, qp0, qpp, q0p)
                re = _border_zeros(*[(xs[k], ys[k], np.real(qs[k])) for k in range(4)])
                im = _border_zeros(*[(xs[k], ys[k], np.imag(qs[k])) for k in range(4)])
                if len(re) == 2 and len(im) == 2:
                    ix, iy, ph = _lines_cross(re[0], re[1], im[0], im[1])
                    if x - dx < ix < x + 2.0 * dx and y - dy < iy < y + 2.0 * dy:
                        found.append([ix, iy, np.round(v)])
            self._detected_vortices = (np.array(found, dtype=cfg.dtype) if found
                                       else np.zeros((0, 3), dtype=cfg.dtype))
            self.solver.vortices_detected = True
        d = self._detected_vortices
        return d[:, 0], d[:, 1], d[:, 2]
