// Host-side line search of the CG minimiser as an OPT-IN alternative to the reference's SciPy call
// (svirl/solvers/cg.py:378-419 minimises the same polynomial with scipy.optimize.minimize(BFGS)).
//
//   G(a_psi, a_A) = sum_ij c[i][j] a_psi^i a_A^j      (c00..c04, c10..c14, c20..c24, c30, c40: cg.h:528-701)
//
// Two variables and a quartic: a damped Newton iteration with the exact Hessian, started at (0, 0) like the
// reference, on the polynomial normalised by max|c| (so the result does not depend on the size of the grid --
// SciPy's BFGS is not scale invariant and runs away from 8192^2 nodes on, profiles/r02_cfg4_adjudicate_8192.json).
// It converges to the local minimum of the basin of (0, 0) to machine precision in 5-10 iterations (~1 us),
// where the Python call costs ~1 ms per CG iteration.  The step differs from SciPy's by its termination error
// (~1e-8 relative), so trajectories agree with the reference to that level, not bit for bit: parity runs keep the
// reference call (cfg.cg_line_search = 'reference', the default).
#include "common.cuh"

namespace {
struct Poly17 {
    double c[5][5];
    void eval(double x, double y, double *f, double g[2], double h[3]) const {
        // value, gradient and Hessian (h = {fxx, fxy, fyy}) by direct summation over the 17 terms
        double xp[5] = {1, x, x * x, x * x * x, x * x * x * x}, yp[5] = {1, y, y * y, y * y * y, y * y * y * y};
        double F = 0, gx = 0, gy = 0, hxx = 0, hxy = 0, hyy = 0;
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) {
                const double cij = c[i][j];
                if (cij == 0.0) continue;
                F += cij * xp[i] * yp[j];
                if (i >= 1) gx += cij * i * xp[i - 1] * yp[j];
                if (j >= 1) gy += cij * j * xp[i] * yp[j - 1];
                if (i >= 2) hxx += cij * i * (i - 1) * xp[i - 2] * yp[j];
                if (i >= 1 && j >= 1) hxy += cij * i * j * xp[i - 1] * yp[j - 1];
                if (j >= 2) hyy += cij * j * (j - 1) * xp[i] * yp[j - 2];
            }
        if (f) *f = F;
        if (g) { g[0] = gx; g[1] = gy; }
        if (h) { h[0] = hxx; h[1] = hxy; h[2] = hyy; }
    }
};
}  // namespace

// c17: the 17 coefficients in the kernel's flat order; alpha_out[2] = (alpha_psi, alpha_A); iters_out optional.
// solveA = 0: c17 holds c0..c4 of the one-variable quartic (cg.h:400-467) and alpha_A = 0.
extern "C" int svl_cg_line_search(const double *c17, int solveA, double *alpha_out, int *iters_out) {
    SVL_REQUIRE(c17 && alpha_out, "null argument");
    Poly17 P;
    memset(&P, 0, sizeof(P));
    if (solveA) {
        for (int j = 0; j < 5; j++) { P.c[0][j] = c17[j]; P.c[1][j] = c17[5 + j]; P.c[2][j] = c17[10 + j]; }
        P.c[3][0] = c17[15]; P.c[4][0] = c17[16];
    } else {
        for (int i = 0; i < 5; i++) P.c[i][0] = c17[i];
    }
    double scale = 0.0;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) scale = fmax(scale, fabs(P.c[i][j]));
    SVL_REQUIRE(scale > 0.0 && isfinite(scale), "line search: coefficients are zero or not finite");
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) P.c[i][j] /= scale;
    // Trust-region Newton: steps never exceed the radius (<= 2 in alpha units, alpha being O(0.1 .. 10) here), so the
    // iteration stays in the basin of (0, 0) although the polynomial is unbounded below (the exponential of the A
    // update is truncated at 4th order).
    double x = 0.0, y = 0.0, f, g[2], h[3], rad = 0.5;
    const double rad_max = 2.0;
    int it = 0;
    for (; it < 500; it++) {
        P.eval(x, y, &f, g, h);
        if (!solveA) { g[1] = 0.0; h[1] = 0.0; h[2] = 1.0; }
        const double gn = fmax(fabs(g[0]), fabs(g[1]));
        if (gn <= 1e-15) break;
        // eigen-decomposition of the symmetric 2x2 Hessian
        const double a = h[0], b = h[1], d = h[2];
        const double mean = 0.5 * (a + d), dif = 0.5 * (a - d), rt = sqrt(dif * dif + b * b);
        const double l1 = mean - rt, l2 = mean + rt;                    // l1 <= l2
        double v1x, v1y;                                                // eigenvector of l1
        if (fabs(b) > 1e-300) { v1x = l1 - d; v1y = b; } else if (a <= d) { v1x = 1; v1y = 0; } else { v1x = 0; v1y = 1; }
        const double vn = sqrt(v1x * v1x + v1y * v1y);
        v1x /= vn; v1y /= vn;
        const double v2x = -v1y, v2y = v1x;
        const double g1 = g[0] * v1x + g[1] * v1y, g2 = g[0] * v2x + g[1] * v2y;
        // s(lam) = -g_i / (l_i + lam); find lam >= max(0, -l1) with |s| <= rad
        auto snorm = [&](double lam) { const double s1 = g1 / (l1 + lam), s2 = g2 / (l2 + lam); return sqrt(s1 * s1 + s2 * s2); };
        double lam = 0.0;
        if (!(l1 > 0.0 && snorm(0.0) <= rad)) {
            double lo = fmax(0.0, -l1), hi = lo + 1.0;
            lo += 1e-14 * fmax(1.0, fabs(l2));
            while (snorm(hi) > rad) hi = lo + 2.0 * (hi - lo) + 1.0;
            if (snorm(lo) <= rad) hi = lo;                              // hard case: interior along the remaining direction
            for (int k = 0; k < 100 && hi > lo; k++) {
                const double mid = 0.5 * (lo + hi);
                if (snorm(mid) > rad) lo = mid; else hi = mid;
                if (hi - lo <= 1e-15 * fmax(1.0, hi)) break;
            }
            lam = hi;
        }
        const double s1 = -g1 / (l1 + lam), s2 = -g2 / (l2 + lam);
        const double sx = s1 * v1x + s2 * v2x, sy = solveA ? s1 * v1y + s2 * v2y : 0.0;
        const double sl = sqrt(sx * sx + sy * sy);
        const double pred = -(g[0] * sx + g[1] * sy + 0.5 * (a * sx * sx + 2.0 * b * sx * sy + d * sy * sy));
        double fn;
        P.eval(x + sx, y + sy, &fn, nullptr, nullptr);
        const double rho = pred > 0.0 ? (f - fn) / pred : -1.0;
        if (rho < 0.25) rad = 0.25 * fmin(rad, sl);
        else if (rho > 0.75 && sl >= 0.99 * rad) rad = fmin(2.0 * rad, rad_max);
        if (rho > 1e-4 || (fn <= f && pred <= 1e-30)) {
            if (x + sx == x && y + sy == y) break;
            x += sx; y += sy;
        }
        if (rad < 1e-17 * (1.0 + fabs(x) + fabs(y))) break;
    }
    alpha_out[0] = x; alpha_out[1] = solveA ? y : 0.0;
    if (iters_out) *iters_out = it;
    return 0;
}
