"""Vectorised sub-cell vortex positions (host side of the detector, svirl/observables/vortex_detector.py:62-145).

The per-candidate arithmetic of ``VortexDetector.vortices`` evaluated on whole arrays of candidate cells
instead of one Python iteration per cell, with the same operand types, the same expression order and the
same NumPy ufuncs -- elementwise IEEE arithmetic does not depend on the array length, and the results are
checked to be bit-identical to the reference's on every fixture (tests/test_triangulate_host.py).  Meant for
grids with millions of vortices (svirl_b200/scale.py), where the interpreter loop is too slow.

All inputs are host arrays indexed [i, j]; `rows0` is the global row of column index 0 of the arrays, so
that a row band of a large grid can be passed."""
import numpy as np


def _edge_zero(x1, y1, f1, x2, y2, f2):
    with np.errstate(all="ignore"):
        return (f2 * x1 - x2 * f1) / (f2 - f1), (f2 * y1 - y2 * f1) / (f2 - f1)


def _cmul(z, w):
    """z * w with separately rounded products, like NumPy's complex SCALAR multiply (the array loop fuses
    multiply-adds, which changes the last bit)."""
    out = np.empty(np.broadcast(z, w).shape, dtype=np.result_type(z, w))
    out.real = z.real * w.real - z.imag * w.imag
    out.imag = z.real * w.imag + z.imag * w.real
    return out


def _two_zeros(xs, ys, fs):
    """Zeros of the linearised f on the four cell edges in the reference's edge order (2-1, 3-2, 4-3, 1-4);
    returns (valid, (x_a, y_a), (x_b, y_b)) with valid = exactly two edges change sign."""
    pairs = ((1, 0), (2, 1), (3, 2), (0, 3))
    hit = np.stack([fs[p] * fs[q] < -1e-10 for p, q in pairs])                   # (4, n)
    zx = np.empty(hit.shape, dtype=np.result_type(xs[0], fs[0]))
    zy = np.empty_like(zx)
    for k, (p, q) in enumerate(pairs):
        zx[k], zy[k] = _edge_zero(xs[p], ys[p], fs[p], xs[q], ys[q], fs[q])
    valid = hit.sum(axis=0) == 2
    order = np.argsort(~hit, axis=0, kind="stable")                              # hit edges first, edge order kept
    cols = np.arange(hit.shape[1])
    first, second = order[0], order[1]
    return valid, (zx[first, cols], zy[first, cols]), (zx[second, cols], zy[second, cols])


def triangulate(cells, psi, a, b, a_ai, b_bi, H, dx, dy, Nxc, rows0=0, dtype=np.float64):
    """(x, y, vorticity) of the candidate cells that pass the reference's tests, ascending cell index.
    cells: flat indices n = i + Nxc*j (global j).  psi, a, b, a_ai, b_bi hold rows rows0 .. of the fields."""
    cells = np.sort(np.asarray(cells, dtype=np.int64))
    i = (cells % Nxc).astype(np.int32)
    jg = (cells // Nxc).astype(np.int32)                 # global row: enters the coordinates
    j = (jg - np.int32(rows0)).astype(np.int32)          # local row: indexes the band
    ip, jp = i + 1, j + 1
    pi = np.pi
    # np.angle at the candidate corners only (elementwise: same values as np.angle(psi)[...])
    t_00, t_p0, t_pp, t_0p = np.angle(psi[i, j]), np.angle(psi[ip, j]), np.angle(psi[ip, jp]), np.angle(psi[i, jp])
    v = - (0.5 / pi) * (
        np.mod(t_p0 - t_00 - dx * a[i, j] + pi, 2.0 * pi)
        + np.mod(t_pp - t_p0 - dy * b[ip, j] + pi, 2.0 * pi)
        + np.mod(t_0p - t_pp + dx * a[i, jp] + pi, 2.0 * pi)
        + np.mod(t_00 - t_0p + dy * b[i, j] + pi, 2.0 * pi)
        - 4.0 * pi
        + dx * dy * H)
    keep = (np.abs(v) > 0.5) & (np.abs(v - np.round(v)) < 0.1)
    i, j, jg, ip, jp, v = i[keep], j[keep], jg[keep], ip[keep], jp[keep], v[keep]
    if i.size == 0:
        z = np.zeros(0, dtype=dtype)
        return z, z.copy(), z.copy()
    x, y = dx * i, dy * jg
    ia00, ia0p = dx * a_ai[i, j], dx * a_ai[i, jp]
    ib00, ibp0 = dy * b_bi[i, j], dy * b_bi[ip, j]
    q00 = psi[i, j]
    qp0 = _cmul(psi[ip, j], np.exp(-1j * (0.75 * ia00 + 0.25 * (ib00 + ia0p - ibp0))))
    qpp = _cmul(psi[ip, jp], np.exp(-1j * (0.5 * (ia00 + ibp0) + 0.5 * (ib00 + ia0p))))
    q0p = _cmul(psi[i, jp], np.exp(-1j * (0.75 * ib00 + 0.25 * (ia00 + ibp0 - ia0p))))
    xs, ys = (x, x + dx, x + dx, x), (y, y, y + dy, y + dy)
    qs = (q00, qp0, qpp, q0p)
    ok_r, r1, r2 = _two_zeros(xs, ys, [np.real(q) for q in qs])
    ok_i, m1, m2 = _two_zeros(xs, ys, [np.imag(q) for q in qs])
    with np.errstate(all="ignore"):
        ux, uy = r1[0] - r2[0], r1[1] - r2[1]
        wx, wy = m1[0] - m2[0], m1[1] - m2[1]
        D = ux * wy - uy * wx
        ph = np.mod(np.abs(np.arctan2(D, uy * wy - ux * wx)), 0.5 * np.pi)
        cp, cq = r1[0] * r2[1] - r1[1] * r2[0], m1[0] * m2[1] - m1[1] * m2[0]
        ix, iy = (cp * wx - ux * cq) / D, (cp * wy - uy * cq) / D
    good = ok_r & ok_i & (np.abs(ph) > 1e-10)
    with np.errstate(invalid="ignore"):
        good &= (x - dx < ix) & (ix < x + 2.0 * dx) & (y - dy < iy) & (iy < y + 2.0 * dy)
    return ix[good].astype(dtype), iy[good].astype(dtype), np.round(v[good]).astype(dtype)
