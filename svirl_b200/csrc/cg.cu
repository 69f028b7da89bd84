// Free energy, Jacobians, line-search coefficients and vector updates of the modified
// nonlinear-CG minimiser.  Reference: svirl/cuda/observables.h:239-362 (energy),
// svirl/cuda/cg.h:5-121 (jacobian psi), :125-301 (jacobian A), :305-474 (5 coefficients),
// :478-731 (17 coefficients), svirl/cuda/utils.h:13-146 (beta sums, axpy/axmy).
// Edge weights follow the "DU" convention everywhere here (svirl/cuda/common.h:13).
#include "common.cuh"

template <typename R> struct Field {   // optional edge field (regular or external potential)
    const R *a, *b;
};

template <typename R>
__device__ __forceinline__ void du_weights(unsigned f, R &wW, R &wE, R &wS, R &wN, R &gw) {
    R mm = (f & NF_MM) ? (R)1 : (R)0, mp = (f & NF_MP) ? (R)1 : (R)0;
    R pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0;
    wW = (R)0.5 * (mm + mp); wE = (R)0.5 * (pm + pp);
    wS = (R)0.5 * (mm + pm); wN = (R)0.5 * (mp + pp);
    gw = (R)0.25 * (wW + wE + wS + wN);
}

// psi1 * U(ph) - psi0   (cg.h:305-311)
template <typename R, typename C>
__device__ __forceinline__ C grad_c(C p0, R s, R c, C p1) {
    C z;
    z.x = p1.x * c + p1.y * s - p0.x;
    z.y = p1.y * c - p1.x * s - p0.y;
    return z;
}

template <typename R>
__device__ __forceinline__ R edge_sum(const R *e, const R *r, size_t n) {
    R p = 0;
    if (e) p += e[n];
    if (r) p += r[n];
    return p;
}

// B - H contribution of one field on cell (i,j): idx*(b[i+1,j]-b[i,j]) - idy*(a[i,j+1]-a[i,j])
template <typename R>
__device__ __forceinline__ R cell_curl(const R *a, const R *b, size_t n, int P, R idx, R idy) {
    return idx * (b[n + 1] - b[n]) - idy * (a[n + P] - a[n]);
}

// ----------------------------------------------------------------------------- energy density of one node
template <typename R, typename C>
__device__ __forceinline__ R node_energy(const Geo &g, int i, int j, size_t n, unsigned f, R kappa2, R eps, R H,
                                         const C *psi, const R *ae, const R *be, const R *a, const R *b) {
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    R e = 0;
    if (f) {
        R wW, wE, wS, wN, gw;
        du_weights<R>(f, wW, wE, wS, wN, gw);
        C p0 = psi[n];
        R p2 = p0.x * p0.x + p0.y * p0.y;
        e += gw * ((R)0.5 * p2 - eps) * p2;
        R s, c;
        if (f & (NF_PM | NF_PP)) {
            sincos_r<R>(dx * edge_sum<R>(ae, a, n), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n + 1]);
            e += wE * idx2 * (z.x * z.x + z.y * z.y);
        }
        if (f & (NF_MP | NF_PP)) {
            sincos_r<R>(dy * edge_sum<R>(be, b, n), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n + g.P]);
            e += wN * idy2 * (z.x * z.x + z.y * z.y);
        }
    }
    if (kappa2 > (R)0 && i < g.Nx - 1 && j < g.Ny - 1) {
        R dB = -H;
        if (ae) dB += cell_curl<R>(ae, be, n, g.P, idx, idy);
        if (a) dB += cell_curl<R>(a, b, n, g.P, idx, idy);
        e += kappa2 * dB * dB;
    }
    return e;
}

template <typename R>
__global__ void __launch_bounds__(256)
k_energy(Geo g, R kappa2, R eps, const R *__restrict__ epsf, R H, const uint8_t *__restrict__ nf,
         const typename V2<R>::type *__restrict__ psi, const R *__restrict__ ae, const R *__restrict__ be,
         const R *__restrict__ a, const R *__restrict__ b, double *partials) {
    typedef typename V2<R>::type C;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    double v[1] = {0.0};
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        v[0] = (double)node_energy<R, C>(g, i, j, n, nf[n], kappa2, epsf ? epsf[n] : eps, H, psi, ae, be, a, b);
    }
    block_sum_to_partials<1>(v, partials, blockIdx.y * gridDim.x + blockIdx.x);
}

// ----------------------------------------------------------------------------- jacobians
template <typename R, typename C>
__device__ __forceinline__ C node_jac_psi(const Geo &g, size_t n, unsigned f, R eps, const C *psi, const R *ae,
                                          const R *be, const R *a, const R *b) {
    const R dx = (R)g.dx, dy = (R)g.dy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    C gj;
    gj.x = 0; gj.y = 0;
    if (f) {
        R wW, wE, wS, wN, gw;
        du_weights<R>(f, wW, wE, wS, wN, gw);
        C p0 = psi[n];
        R p = p0.x * p0.x + p0.y * p0.y - eps;
        gj.x += (R)2.0 * gw * p * p0.x;
        gj.y += (R)2.0 * gw * p * p0.y;
        R s, c;
        const int P = g.P;
        // g_grad_jac_psi(psi0, ph, psi1) = 2 (psi0 - psi1 U(ph))   (cg.h:5-12)
        if (f & (NF_MM | NF_MP)) {
            sincos_r<R>(-dx * edge_sum<R>(ae, a, n - 1), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n - 1]);
            gj.x += wW * idx2 * ((R)-2.0 * z.x); gj.y += wW * idx2 * ((R)-2.0 * z.y);
        }
        if (f & (NF_PM | NF_PP)) {
            sincos_r<R>(dx * edge_sum<R>(ae, a, n), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n + 1]);
            gj.x += wE * idx2 * ((R)-2.0 * z.x); gj.y += wE * idx2 * ((R)-2.0 * z.y);
        }
        if (f & (NF_MM | NF_PM)) {
            sincos_r<R>(-dy * edge_sum<R>(be, b, n - P), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n - P]);
            gj.x += wS * idy2 * ((R)-2.0 * z.x); gj.y += wS * idy2 * ((R)-2.0 * z.y);
        }
        if (f & (NF_MP | NF_PP)) {
            sincos_r<R>(dy * edge_sum<R>(be, b, n), &s, &c);
            C z = grad_c<R, C>(p0, s, c, psi[n + P]);
            gj.x += wN * idy2 * ((R)-2.0 * z.x); gj.y += wN * idy2 * ((R)-2.0 * z.y);
        }
    }
    R dxdy = dx * dy;
    gj.x *= dxdy; gj.y *= dxdy;
    return gj;
}

template <typename R>
__global__ void __launch_bounds__(256)
k_jac_psi(Geo g, R eps, const R *__restrict__ epsf, const uint8_t *__restrict__ nf,
          const typename V2<R>::type *__restrict__ psi, const R *__restrict__ ae, const R *__restrict__ be,
          const R *__restrict__ a, const R *__restrict__ b, typename V2<R>::type *__restrict__ out) {
    typedef typename V2<R>::type C;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        out[n] = node_jac_psi<R, C>(g, n, nf[n], epsf ? epsf[n] : eps, psi, ae, be, a, b);
    }
}

// magnetic ("curl curl") part on the a-edge / b-edge of node (i,j) before the kappa2 factor
// (cg.h:176-217, 240-282; the same stencil as current_density, observables.h:66-155): quirk Q10.
template <typename R>
__device__ __forceinline__ R curlcurl_a(const Geo &g, int i, int j, size_t n, R H, const R *ae, const R *be,
                                        const R *a, const R *b) {
    const R idy = (R)g.idy, idy2 = (R)g.idy2, idxy = (R)g.idxy;
    const int P = g.P;
    R v = 0, dd = 1;
    if (j == 0) { v -= (R)2.0 * H * idy; dd = 2; }
    else if (j + 1 == g.Ny) { v += (R)2.0 * H * idy; dd = 2; }
    if (ae) v += (R)2.0 / dd * idy2 * ae[n];
    if (a) v += (R)2.0 * idy2 * a[n];
    if (j > 0) {
        if (ae) v += (-idy2 * ae[n - P] + idxy * be[n - P] - idxy * be[n - P + 1]);
        if (a) v += dd * (-idy2 * a[n - P] + idxy * b[n - P] - idxy * b[n - P + 1]);
    }
    if (j + 1 < g.Ny) {
        if (ae) v += (-idy2 * ae[n + P] - idxy * be[n] + idxy * be[n + 1]);
        if (a) v += dd * (-idy2 * a[n + P] - idxy * b[n] + idxy * b[n + 1]);
    }
    return v;
}

template <typename R>
__device__ __forceinline__ R curlcurl_b(const Geo &g, int i, int j, size_t n, R H, const R *ae, const R *be,
                                        const R *a, const R *b) {
    const R idx = (R)g.idx, idx2 = (R)g.idx2, idxy = (R)g.idxy;
    const int P = g.P;
    R v = 0, dd = 1;
    if (i == 0) { v += (R)2.0 * H * idx; dd = 2; }
    else if (i + 1 == g.Nx) { v -= (R)2.0 * H * idx; dd = 2; }
    if (be) v += (R)2.0 / dd * idx2 * be[n];
    if (b) v += (R)2.0 * idx2 * b[n];
    if (i > 0) {
        if (ae) v += (-idx2 * be[n - 1] + idxy * ae[n - 1] - idxy * ae[n - 1 + P]);
        if (a) v += dd * (-idx2 * b[n - 1] + idxy * a[n - 1] - idxy * a[n - 1 + P]);
    }
    if (i + 1 < g.Nx) {
        if (ae) v += (-idx2 * be[n + 1] - idxy * ae[n] + idxy * ae[n + P]);
        if (a) v += dd * (-idx2 * b[n + 1] - idxy * a[n] + idxy * a[n + P]);
    }
    return v;
}

template <typename R>
__global__ void __launch_bounds__(256)
k_jac_A(Geo g, R kappa2, R H, const uint8_t *__restrict__ nf, const typename V2<R>::type *__restrict__ psi,
        const R *__restrict__ ae, const R *__restrict__ be, const R *__restrict__ a, const R *__restrict__ b,
        R *__restrict__ oa, R *__restrict__ ob) {
    typedef typename V2<R>::type C;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Nx || j >= g.j1) return;
    size_t n = g.at(i, j);
    unsigned f = nf[n];
    C p0 = psi[n];
    R mp = (f & NF_MP) ? (R)1 : (R)0, pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0;
    if (i < g.Nx - 1) {
        R v = kappa2 * curlcurl_a<R>(g, i, j, n, H, ae, be, a, b);
        if (f & (NF_PM | NF_PP))
            v += -((R)0.5 * (pm + pp)) * idx * js_link<R, C>(p0, dx * edge_sum<R>(ae, a, n), psi[n + 1]);
        oa[n] = (R)2.0 * dx * dy * v;
    }
    if (j < g.Ny - 1) {
        R v = kappa2 * curlcurl_b<R>(g, i, j, n, H, ae, be, a, b);
        if (f & (NF_MP | NF_PP))
            v += -((R)0.5 * (mp + pp)) * idy * js_link<R, C>(p0, dy * edge_sum<R>(be, b, n), psi[n + g.P]);
        ob[n] = (R)2.0 * dx * dy * v;
    }
}

// ----------------------------------------------------------------------------- line-search coefficients
// NV = 5: c0..c4 (cg.h:400-467; magnetic term from the regular potential only, cg.h:449-455).
// NV = 17: c00..c04, c10..c14, c20..c24, c30, c40 (cg.h:528-701).
// Quirk Q11: the spatial linear coefficient is NOT used by these kernels (scalar eps only).
template <typename R, int NV>
__global__ void __launch_bounds__(256)
k_coef(Geo g, R kappa2, R eps, R H, const uint8_t *__restrict__ nf, const typename V2<R>::type *__restrict__ psi,
       const typename V2<R>::type *__restrict__ dpsi, const R *__restrict__ ae, const R *__restrict__ be,
       const R *__restrict__ a, const R *__restrict__ b, const R *__restrict__ da, const R *__restrict__ db,
       double *partials) {
    typedef typename V2<R>::type C;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    // w is 0, 1/2 or 1, so (-w*i2)/3 == -w*(i2/3) bit for bit: the divisions of cg.h:600-640 leave the node loop
    const R idx2_3 = idx2 / (R)3.0, idy2_3 = idy2 / (R)3.0, idx2_12 = idx2 / (R)12.0, idy2_12 = idy2 / (R)12.0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = 0.0;
    // index helpers: c[r][k] -> flat
    const int C0 = 0, C1 = NV == 17 ? 5 : 1, C2 = NV == 17 ? 10 : 2, C3 = NV == 17 ? 15 : 3, C4 = NV == 17 ? 16 : 4;
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        unsigned f = nf[n];
        if (f) {
            R wW, wE, wS, wN, gw;
            du_weights<R>(f, wW, wE, wS, wN, gw);
            C p0 = psi[n], d0 = dpsi[n];
            R p2 = p0.x * p0.x + p0.y * p0.y, d2 = d0.x * d0.x + d0.y * d0.y;
            R tw = (R)2.0 * (p0.x * d0.x + p0.y * d0.y);
            v[C0] += (double)(gw * ((R)0.5 * p2 - eps) * p2);
            v[C1] += (double)(gw * tw * (p2 - eps));
            v[C2] += (double)(gw * (-eps * d2 + (R)0.5 * tw * tw + p2 * d2));
            v[C3] += (double)(gw * tw * d2);
            v[C4] += (double)(gw * (R)0.5 * d2 * d2);
#pragma unroll
            for (int dir = 0; dir < 2; dir++) {
                bool on = dir == 0 ? (f & (NF_PM | NF_PP)) : (f & (NF_MP | NF_PP));
                if (!on) continue;
                size_t n1 = dir == 0 ? n + 1 : n + g.P;
                R w = dir == 0 ? wE : wN, i2 = dir == 0 ? idx2 : idy2, d = dir == 0 ? dx : dy;
                const R i2_3 = dir == 0 ? idx2_3 : idy2_3, i2_12 = dir == 0 ? idx2_12 : idy2_12;   // Taylor 1/3, 1/12 folded in
                R ph = 0;
                if (dir == 0) { if (ae) ph += d * ae[n]; if (a) ph += d * a[n]; }
                else { if (be) ph += d * be[n]; if (b) ph += d * b[n]; }
                R s, c;
                sincos_r<R>(ph, &s, &c);
                C p1 = psi[n1], d1 = dpsi[n1];
                C zp = grad_c<R, C>(p0, s, c, p1), zd = grad_c<R, C>(d0, s, c, d1);
                v[C0] += (double)(w * i2 * (zp.x * zp.x + zp.y * zp.y));
                v[C1] += (double)(w * i2 * (R)2.0 * (zp.x * zd.x + zp.y * zd.y));   // Re(conj(zp) zd)
                v[C2] += (double)(w * i2 * (zd.x * zd.x + zd.y * zd.y));
                if (NV == 17) {
                    R dph = d * (dir == 0 ? da[n] : db[n]);
                    R dph2 = dph * dph;
                    // z = x0 * U(-ph) * conj(x1), U(-ph) = c + i s
#define ZMUL(x0, x1, zr, zi)                                             \
    {                                                                    \
        R ur = x0.x * c - x0.y * s, ui = x0.x * s + x0.y * c;            \
        zr = ur * x1.x + ui * x1.y;                                      \
        zi = ui * x1.x - ur * x1.y;                                      \
    }
                    R zr, zi, z2r, z2i;
                    ZMUL(p0, p1, zr, zi);
                    v[1] += (double)(w * i2 * (R)2.0 * zi * dph);
                    v[2] += (double)(w * i2 * zr * dph2);
                    v[3] += (double)(-w * i2_3 * zi * dph2 * dph);
                    v[4] += (double)(-w * i2_12 * zr * dph2 * dph2);
                    ZMUL(p0, d1, zr, zi);
                    ZMUL(d0, p1, z2r, z2i);
                    zr += z2r; zi += z2i;
                    v[6] += (double)(w * i2 * (R)2.0 * zi * dph);
                    v[7] += (double)(w * i2 * zr * dph2);
                    v[8] += (double)(-w * i2_3 * zi * dph2 * dph);
                    v[9] += (double)(-w * i2_12 * zr * dph2 * dph2);
                    ZMUL(d0, d1, zr, zi);
                    v[11] += (double)(w * i2 * (R)2.0 * zi * dph);
                    v[12] += (double)(w * i2 * zr * dph2);
                    v[13] += (double)(-w * i2_3 * zi * dph2 * dph);
                    v[14] += (double)(-w * i2_12 * zr * dph2 * dph2);
#undef ZMUL
                }
            }
        }
        if (kappa2 > (R)0 && i < g.Nx - 1 && j < g.Ny - 1) {
            if (NV == 17) {
                R BH = -H;
                if (ae) BH += cell_curl<R>(ae, be, n, g.P, idx, idy);
                if (a) BH += cell_curl<R>(a, b, n, g.P, idx, idy);
                R dB = cell_curl<R>(da, db, n, g.P, idx, idy);
                v[0] += (double)(kappa2 * BH * BH);
                v[1] += (double)(kappa2 * (R)2.0 * BH * dB);
                v[2] += (double)(kappa2 * dB * dB);
            } else {
                R dB = cell_curl<R>(a, b, n, g.P, idx, idy) - H;
                v[0] += (double)(kappa2 * dB * dB);
            }
        }
    }
    block_sum_to_partials<NV>(v, partials, blockIdx.y * gridDim.x + blockIdx.x);
}

// ----------------------------------------------------------------------------- vector kernels
// z = alpha*x + sgn*y over all plane elements (padding is zero and stays zero).
template <typename R>
__global__ void __launch_bounds__(256) k_axy(const R *x, const R *y, R *z, R alpha, R sgn, size_t n) {   // z may alias x or y
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        z[i] = alpha * x[i] + sgn * y[i];
}

// partial sums of g.(g - gp) and gp.gp over plane elements
template <typename R>
__global__ void __launch_bounds__(256) k_beta_sums(const R *__restrict__ gj, const R *__restrict__ gp, size_t n, double *partials) {
    double v[2] = {0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        R x = gj[i], y = gp[i];
        v[0] += (double)(x * (x - y));
        v[1] += (double)(y * y);
    }
    block_sum_to_partials<2>(v, partials, blockIdx.x);
}

// ----------------------------------------------------------------------------- fused CG kernels
// Both Jacobians of node (i,j) (psi, its a-edge and its b-edge) in one pass, plus the four
// Polak-Ribiere partial sums  sum g.(g-gp), sum gp.gp  for psi and for A (svirl/cuda/utils.h:13-70).
template <typename R, bool SOLVEA, bool PREV>
__global__ void __launch_bounds__(256)
k_cg_grad(Geo g, R kappa2, R eps, const R *__restrict__ epsf, R H, const uint8_t *__restrict__ nf,
          const typename V2<R>::type *__restrict__ psi, const R *__restrict__ ae, const R *__restrict__ be,
          const R *__restrict__ a, const R *__restrict__ b, typename V2<R>::type *__restrict__ gpsi,
          R *__restrict__ ga, R *__restrict__ gb, const typename V2<R>::type *__restrict__ ppsi,
          const R *__restrict__ pa, const R *__restrict__ pb, double *partials) {
    typedef typename V2<R>::type C;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < g.Nx && j < g.j1) {
        size_t n = g.at(i, j);
        unsigned f = nf[n];
        C gj = node_jac_psi<R, C>(g, n, f, epsf ? epsf[n] : eps, psi, ae, be, a, b);
        gpsi[n] = gj;
        if (PREV) {
            C p = ppsi[n];
            v[0] = (double)(gj.x * (gj.x - p.x) + gj.y * (gj.y - p.y));
            v[1] = (double)(p.x * p.x + p.y * p.y);
        }
        if (SOLVEA) {
            C p0 = psi[n];
            R mp = (f & NF_MP) ? (R)1 : (R)0, pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0;
            if (i < g.Nx - 1) {
                R w = kappa2 * curlcurl_a<R>(g, i, j, n, H, ae, be, a, b);
                if (f & (NF_PM | NF_PP))
                    w += -((R)0.5 * (pm + pp)) * idx * js_link<R, C>(p0, dx * edge_sum<R>(ae, a, n), psi[n + 1]);
                w = (R)2.0 * dx * dy * w;
                ga[n] = w;
                if (PREV) { R p = pa[n]; v[2] += (double)(w * (w - p)); v[3] += (double)(p * p); }
            }
            if (j < g.Ny - 1) {
                R w = kappa2 * curlcurl_b<R>(g, i, j, n, H, ae, be, a, b);
                if (f & (NF_MP | NF_PP))
                    w += -((R)0.5 * (mp + pp)) * idy * js_link<R, C>(p0, dy * edge_sum<R>(be, b, n), psi[n + g.P]);
                w = (R)2.0 * dx * dy * w;
                gb[n] = w;
                if (PREV) { R p = pb[n]; v[2] += (double)(w * (w - p)); v[3] += (double)(p * p); }
            }
        }
    }
    if (PREV) block_sum_to_partials<4>(v, partials, blockIdx.y * gridDim.x + blockIdx.x);
}

// beta = max(num/den, 0) in real_t, nan -> 0 (divide_scalars_positive, utils.h:140-146); stays on the device
template <typename R>
__global__ void k_beta_from_sums(const double *__restrict__ sums, double *beta) {
    if (threadIdx.x < 2) {
        R q = (R)sums[2 * threadIdx.x] / (R)sums[2 * threadIdx.x + 1];
        beta[threadIdx.x] = (q > (R)0) ? (double)q : 0.0;
    }
}

// z = alpha*x + sgn*y on up to three planes in one launch; alpha either by value or read from
// device memory (beta[which]), like the reference's axmy kernels (utils.h:97-114).
template <typename R>
struct Axy3 { const R *x[3]; const R *y[3]; R *z[3]; size_t n[3]; int which[3]; };
template <typename R>
__global__ void __launch_bounds__(256) k_axy3(Axy3<R> d, R alpha0, R alpha1, const double *__restrict__ dev_alpha, R sgn) {
    const int pl = blockIdx.y;
    if (!d.x[pl]) return;
    R al = d.which[pl] ? alpha1 : alpha0;
    if (dev_alpha) al = (R)dev_alpha[d.which[pl]];
    const R *x = d.x[pl], *y = d.y[pl];
    R *z = d.z[pl];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < d.n[pl]; i += (size_t)gridDim.x * 256)
        z[i] = al * x[i] + sgn * y[i];
}

// ----------------------------------------------------------------------------- host wrappers
#define GRID2D(c) dim3 bdim(32, 8), gdim(((c)->g.Nx + 31) / 32, ((c)->g.j1 - (c)->g.j0 + 7) / 8)
#define EDGE_A(buf, R) ((buf) ? (const R *)(buf)->p[0] : nullptr)
#define EDGE_B(buf, R) ((buf) ? (const R *)(buf)->p[1] : nullptr)

// slab_ok: the entry point adds its sums over all ranks (free energy, fused CG iteration: validated on 2-8 GPUs with
// tests/slab_gpu_check.py, SLAB_CG=4); the single coefficient kernels produce per-context sums and refuse row slabs
static int check_state(const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, const svl_buf *epsf, bool slab_ok = false) {
    SVL_REQUIRE(psi && psi->kind == SVL_NODE_C, "psi must be SVL_NODE_C");
    SVL_REQUIRE(slab_ok || !(psi->ctx && psi->ctx->slab_on),
                "this kernel-level entry point sums over one context only: on row slabs use svl_free_energy / svl_cg_begin / svl_cg_end");
    SVL_REQUIRE(!ab || ab->kind == SVL_EDGE, "ab must be SVL_EDGE");
    SVL_REQUIRE(!abei || abei->kind == SVL_EDGE, "abei must be SVL_EDGE");
    SVL_REQUIRE(!epsf || epsf->kind == SVL_NODE_R, "eps_field must be SVL_NODE_R");
    return 0;
}

template <typename R>
static int energy_t(svl_ctx *c, double kappa2, double eps, const svl_buf *epsf, double H, const svl_buf *psi,
                    const svl_buf *abei, const svl_buf *ab, double *E) {
    typedef typename V2<R>::type C;
    GRID2D(c);
    int nb = gdim.x * gdim.y;
    SVL_TRY(svl_ensure_partials(c, nb));
    k_energy<R><<<gdim, bdim, 0, c->stream>>>(c->g, (R)kappa2, (R)eps, epsf ? (const R *)epsf->p[0] : nullptr, (R)H, c->nf,
                                             (const C *)psi->p[0], EDGE_A(abei, R), EDGE_B(abei, R), EDGE_A(ab, R),
                                             EDGE_B(ab, R), c->partials);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    if (c->slab_on) {      // row slabs: the per-rank sums are added over the residual board
        SVL_TRY(svl_finish_sum(c, nb, 1, (double)((R)c->g.dx * (R)c->g.dy), nullptr));
        SVL_TRY(svl_board_allsum(c, c->d_result, 1));
        SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        *E = c->h_result[0];
        return 0;
    }
    return svl_finish_sum(c, nb, 1, (double)((R)c->g.dx * (R)c->g.dy), E);
}

extern "C" int svl_free_energy(svl_ctx *c, double kappa2, double eps, const svl_buf *epsf, double H, const svl_buf *psi,
                               const svl_buf *abei, const svl_buf *ab, double *E) {
    SVL_REQUIRE(c && E, "null argument");
    SVL_TRY(check_state(psi, abei, ab, epsf, true));
    if (c->rsize == 4) return energy_t<float>(c, kappa2, eps, epsf, H, psi, abei, ab, E);
    return energy_t<double>(c, kappa2, eps, epsf, H, psi, abei, ab, E);
}

template <typename R>
static int jac_psi_t(svl_ctx *c, double eps, const svl_buf *epsf, const svl_buf *psi, const svl_buf *abei,
                     const svl_buf *ab, svl_buf *out) {
    typedef typename V2<R>::type C;
    GRID2D(c);
    k_jac_psi<R><<<gdim, bdim, 0, c->stream>>>(c->g, (R)eps, epsf ? (const R *)epsf->p[0] : nullptr, c->nf,
                                              (const C *)psi->p[0], EDGE_A(abei, R), EDGE_B(abei, R), EDGE_A(ab, R),
                                              EDGE_B(ab, R), (C *)out->p[0]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

extern "C" int svl_jacobian_psi(svl_ctx *c, double kappa2, double eps, const svl_buf *epsf, double H, const svl_buf *psi,
                                const svl_buf *abei, const svl_buf *ab, svl_buf *out) {
    (void)kappa2; (void)H;
    SVL_REQUIRE(c, "null context");
    SVL_TRY(check_state(psi, abei, ab, epsf));
    SVL_REQUIRE(out && out->kind == SVL_NODE_C && out != psi, "out must be a distinct SVL_NODE_C buffer");
    if (c->rsize == 4) return jac_psi_t<float>(c, eps, epsf, psi, abei, ab, out);
    return jac_psi_t<double>(c, eps, epsf, psi, abei, ab, out);
}

template <typename R>
static int jac_A_t(svl_ctx *c, double kappa2, double H, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab,
                   svl_buf *out) {
    typedef typename V2<R>::type C;
    GRID2D(c);
    k_jac_A<R><<<gdim, bdim, 0, c->stream>>>(c->g, (R)kappa2, (R)H, c->nf, (const C *)psi->p[0], EDGE_A(abei, R),
                                            EDGE_B(abei, R), EDGE_A(ab, R), EDGE_B(ab, R), (R *)out->p[0], (R *)out->p[1]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

extern "C" int svl_jacobian_A(svl_ctx *c, double kappa2, double H, const svl_buf *psi, const svl_buf *abei,
                              const svl_buf *ab, svl_buf *out) {
    SVL_REQUIRE(c, "null context");
    SVL_TRY(check_state(psi, abei, ab, nullptr));
    SVL_REQUIRE(out && out->kind == SVL_EDGE && out != ab, "out must be a distinct SVL_EDGE buffer");
    if (c->rsize == 4) return jac_A_t<float>(c, kappa2, H, psi, abei, ab, out);
    return jac_A_t<double>(c, kappa2, H, psi, abei, ab, out);
}

template <typename R, int NV>
static int coef_t(svl_ctx *c, double kappa2, double eps, double H, const svl_buf *psi, const svl_buf *dpsi,
                  const svl_buf *abei, const svl_buf *ab, const svl_buf *dab, double *out) {
    typedef typename V2<R>::type C;
    GRID2D(c);
    int nb = gdim.x * gdim.y;
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * NV));
    k_coef<R, NV><<<gdim, bdim, 0, c->stream>>>(c->g, (R)kappa2, (R)eps, (R)H, c->nf, (const C *)psi->p[0],
                                               (const C *)dpsi->p[0], EDGE_A(abei, R), EDGE_B(abei, R), EDGE_A(ab, R),
                                               EDGE_B(ab, R), EDGE_A(dab, R), EDGE_B(dab, R), c->partials);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return svl_finish_sum(c, nb, NV, (double)((R)c->g.dx * (R)c->g.dy), out);
}

extern "C" int svl_cg_coef_psi(svl_ctx *c, double kappa2, double eps, double H, const svl_buf *psi, const svl_buf *dpsi,
                               const svl_buf *abei, const svl_buf *ab, double *c5) {
    SVL_REQUIRE(c && c5, "null argument");
    SVL_TRY(check_state(psi, abei, ab, nullptr));
    SVL_REQUIRE(dpsi && dpsi->kind == SVL_NODE_C, "dpsi must be SVL_NODE_C");
    SVL_REQUIRE(!(kappa2 > 0) || ab, "finite kappa needs ab");
    if (c->rsize == 4) return coef_t<float, 5>(c, kappa2, eps, H, psi, dpsi, abei, ab, nullptr, c5);
    return coef_t<double, 5>(c, kappa2, eps, H, psi, dpsi, abei, ab, nullptr, c5);
}

extern "C" int svl_cg_coef(svl_ctx *c, double kappa2, double eps, double H, const svl_buf *psi, const svl_buf *dpsi,
                           const svl_buf *abei, const svl_buf *ab, const svl_buf *dab, double *c17) {
    SVL_REQUIRE(c && c17, "null argument");
    SVL_TRY(check_state(psi, abei, ab, nullptr));
    SVL_REQUIRE(dpsi && dpsi->kind == SVL_NODE_C, "dpsi must be SVL_NODE_C");
    SVL_REQUIRE(dab && dab->kind == SVL_EDGE, "dab must be SVL_EDGE");
    if (c->rsize == 4) return coef_t<float, 17>(c, kappa2, eps, H, psi, dpsi, abei, ab, dab, c17);
    return coef_t<double, 17>(c, kappa2, eps, H, psi, dpsi, abei, ab, dab, c17);
}

template <typename R>
static int beta_sums_t(svl_ctx *c, const svl_buf *gj, const svl_buf *gp, double *sums /* [2] */) {
    int nreal = gj->kind == SVL_NODE_C ? 2 : 1;
    size_t n = (size_t)c->g.rows * c->g.P * nreal;
    int nb = svl_nblocks(n, 256 * 8);
    if (nb > 1184) nb = 1184;
    double tot[2] = {0.0, 0.0};
    for (int k = 0; k < 2; k++) {
        if (!gj->bytes[k]) continue;
        SVL_TRY(svl_ensure_partials(c, (size_t)nb * 2));
        // only owned rows contribute: start at plane row SVL_HALO
        size_t off = (size_t)SVL_HALO * c->g.P * nreal, cnt = (size_t)(c->g.j1 - c->g.j0) * c->g.P * nreal;
        k_beta_sums<R><<<nb, 256, 0, c->stream>>>((const R *)gj->p[k] + off, (const R *)gp->p[k] + off, cnt, c->partials);
        SVL_CHECK(cudaGetLastError());
        c->stat_launches += 1;
        double s[2];
        SVL_TRY(svl_finish_sum(c, nb, 2, 1.0, s));
        tot[0] += s[0]; tot[1] += s[1];
    }
    sums[0] = tot[0]; sums[1] = tot[1];
    return 0;
}

extern "C" int svl_cg_beta(svl_ctx *c, const svl_buf *gj, const svl_buf *gp, double *beta) {
    SVL_REQUIRE(c && gj && gp && beta, "null argument");
    SVL_REQUIRE(gj->kind == gp->kind && (gj->kind == SVL_NODE_C || gj->kind == SVL_EDGE), "beta needs NODE_C or EDGE buffers");
    double s[2];
    if (c->rsize == 4) SVL_TRY(beta_sums_t<float>(c, gj, gp, s));
    else SVL_TRY(beta_sums_t<double>(c, gj, gp, s));
    // divide_scalars_positive (utils.h:140-146): max(num/den, 0) in real_t; fmax(nan, 0) = 0
    double q;
    if (c->rsize == 4) q = (double)((float)s[0] / (float)s[1]);
    else q = s[0] / s[1];
    *beta = (q > 0.0) ? q : 0.0;
    return 0;
}

template <typename R>
static int axy_t(svl_ctx *c, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha, double sgn) {
    int nreal = x->kind == SVL_NODE_C ? 2 : 1;
    size_t n = (size_t)c->g.rows * c->g.P * nreal;
    int nb = svl_nblocks(n, 256 * 4);
    if (nb > 148 * 16) nb = 148 * 16;
    for (int k = 0; k < 2; k++) {
        if (!x->bytes[k]) continue;
        k_axy<R><<<nb, 256, 0, c->stream>>>((const R *)x->p[k], (const R *)y->p[k], (R *)z->p[k], (R)alpha, (R)sgn, n);
        SVL_CHECK(cudaGetLastError());
        c->stat_launches += 1;
    }
    return 0;
}

static int axy(svl_ctx *c, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha, double sgn) {
    SVL_REQUIRE(c && x && y && z, "null argument");
    SVL_REQUIRE(x->kind == y->kind && x->kind == z->kind && (x->kind == SVL_NODE_C || x->kind == SVL_EDGE),
                "axpy/axmy need three NODE_C or three EDGE buffers");
    if (c->rsize == 4) return axy_t<float>(c, x, y, z, alpha, sgn);
    return axy_t<double>(c, x, y, z, alpha, sgn);
}

extern "C" int svl_axmy(svl_ctx *c, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha) {
    return axy(c, x, y, z, alpha, -1.0);
}
extern "C" int svl_axpy(svl_ctx *c, const svl_buf *x, const svl_buf *y, svl_buf *z, double alpha) {
    return axy(c, x, y, z, alpha, 1.0);
}

// ----------------------------------------------------------------------------- fused CG iteration halves
// cg_fused.cu: three-pass iteration (default); the composition below is kept as a cross-check
// (option "cg_fused" = 0)
int svl_cgf_begin(svl_ctx *c, int solveA, int have_prev, double kappa2, double eps, const svl_buf *epsf, double H,
                  const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi, svl_buf *g_psi_prev,
                  svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A, double *beta, double *c_out);
int svl_cgf_end(svl_ctx *c, int solveA, double kappa2, double eps, const svl_buf *epsf, double H, svl_buf *psi,
                const svl_buf *abei, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi,
                double alpha_A, double *E_out);

template <typename R>
static int axy3_t(svl_ctx *c, const svl_buf *xp, const svl_buf *yp, svl_buf *zp, const svl_buf *xA, const svl_buf *yA,
                  svl_buf *zA, double a0, double a1, const double *dev_alpha, double sgn) {
    Axy3<R> d;
    memset(&d, 0, sizeof(d));
    size_t plane = (size_t)c->g.rows * c->g.P;
    d.x[0] = (const R *)xp->p[0]; d.y[0] = (const R *)yp->p[0]; d.z[0] = (R *)zp->p[0]; d.n[0] = 2 * plane; d.which[0] = 0;
    if (xA) {
        for (int k = 0; k < 2; k++) {
            d.x[1 + k] = (const R *)xA->p[k]; d.y[1 + k] = (const R *)yA->p[k]; d.z[1 + k] = (R *)zA->p[k];
            d.n[1 + k] = plane; d.which[1 + k] = 1;
        }
    }
    int nb = svl_nblocks(2 * plane, 256 * 8);
    if (nb > 148 * 8) nb = 148 * 8;
    k_axy3<R><<<dim3(nb, xA ? 3 : 1), 256, 0, c->stream>>>(d, (R)a0, (R)a1, dev_alpha, (R)sgn);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

template <typename R>
static int cg_begin_t(svl_ctx *c, int solveA, int have_prev, double kappa2, double eps, const svl_buf *epsf, double H,
                      const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi, svl_buf *g_psi_prev,
                      svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A, double *beta, double *c_out) {
    typedef typename V2<R>::type C;
    GRID2D(c);
    int nb = gdim.x * gdim.y;
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * 17));
    const R *epf = epsf ? (const R *)epsf->p[0] : nullptr;
#define GRAD_ARGS c->g, (R)kappa2, (R)eps, epf, (R)H, c->nf, (const C *)psi->p[0], EDGE_A(abei, R), EDGE_B(abei, R), \
                  EDGE_A(ab, R), EDGE_B(ab, R), (C *)g_psi->p[0], solveA ? (R *)g_A->p[0] : nullptr,                 \
                  solveA ? (R *)g_A->p[1] : nullptr, (const C *)g_psi_prev->p[0],                                    \
                  solveA ? (const R *)g_A_prev->p[0] : nullptr, solveA ? (const R *)g_A_prev->p[1] : nullptr, c->partials
    if (solveA && have_prev) k_cg_grad<R, true, true><<<gdim, bdim, 0, c->stream>>>(GRAD_ARGS);
    else if (solveA) k_cg_grad<R, true, false><<<gdim, bdim, 0, c->stream>>>(GRAD_ARGS);
    else if (have_prev) k_cg_grad<R, false, true><<<gdim, bdim, 0, c->stream>>>(GRAD_ARGS);
    else k_cg_grad<R, false, false><<<gdim, bdim, 0, c->stream>>>(GRAD_ARGS);
#undef GRAD_ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    double *dbeta = c->d_result + 32;                 // device-resident beta[2]
    if (have_prev) {
        SVL_TRY(svl_finish_sum(c, nb, 4, 1.0, nullptr));          // d_result[0..3], no host read
        k_beta_from_sums<R><<<1, 32, 0, c->stream>>>(c->d_result, dbeta);
        SVL_CHECK(cudaGetLastError());
        if (!solveA) { /* beta[1] unused */ }
    } else {
        // first iteration of a cg() call: keep the betas of the previous call (quirk Q6)
        SVL_CHECK(cudaMemcpyAsync(dbeta, beta, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    // direction update d <- beta*d - g with beta read on the device
    SVL_TRY(axy3_t<R>(c, d_psi, g_psi, d_psi, solveA ? d_A : nullptr, g_A, d_A, 0.0, 0.0, dbeta, -1.0));
    // Quirk Q11: the coefficient kernels use the scalar eps (0.0 when eps is a field)
    double eps_coef = epsf ? 0.0 : eps;
    int rc;
    if (solveA) rc = svl_cg_coef(c, kappa2, eps_coef, H, psi, d_psi, abei, ab, d_A, c_out);   // one host sync
    else rc = svl_cg_coef_psi(c, kappa2, eps_coef, H, psi, d_psi, abei, ab, c_out);
    if (rc) return rc;
    SVL_CHECK(cudaMemcpyAsync(c->h_result + 32, dbeta, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    beta[0] = c->h_result[32];
    if (solveA) beta[1] = c->h_result[33];
    return 0;
}

extern "C" int svl_cg_begin(svl_ctx *c, int solveA, int have_prev, double kappa2, double eps, const svl_buf *epsf,
                            double H, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi,
                            svl_buf *g_psi_prev, svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A,
                            double *beta, double *c_out) {
    SVL_REQUIRE(c && beta && c_out, "null argument");
    SVL_TRY(check_state(psi, abei, ab, epsf, true));
    SVL_REQUIRE(g_psi && g_psi_prev && d_psi && g_psi->kind == SVL_NODE_C && g_psi_prev->kind == SVL_NODE_C &&
                d_psi->kind == SVL_NODE_C, "psi-side CG buffers must be SVL_NODE_C");
    SVL_REQUIRE(!solveA || (g_A && g_A_prev && d_A && ab && g_A->kind == SVL_EDGE && g_A_prev->kind == SVL_EDGE &&
                            d_A->kind == SVL_EDGE), "A-side CG buffers must be SVL_EDGE");
    SVL_REQUIRE(c->opt_cg_fused || !c->slab_on, "CG on slabs needs the fused iteration (option cg_fused = 1)");
    if (c->opt_cg_fused)
        return svl_cgf_begin(c, solveA, have_prev, kappa2, eps, epsf, H, psi, abei, ab, g_psi, g_psi_prev, d_psi, g_A,
                             g_A_prev, d_A, beta, c_out);
    if (c->rsize == 4) return cg_begin_t<float>(c, solveA, have_prev, kappa2, eps, epsf, H, psi, abei, ab, g_psi, g_psi_prev,
                                                d_psi, g_A, g_A_prev, d_A, beta, c_out);
    return cg_begin_t<double>(c, solveA, have_prev, kappa2, eps, epsf, H, psi, abei, ab, g_psi, g_psi_prev, d_psi, g_A,
                              g_A_prev, d_A, beta, c_out);
}

extern "C" int svl_cg_end(svl_ctx *c, int solveA, double kappa2, double eps, const svl_buf *epsf, double H, svl_buf *psi,
                          const svl_buf *abei, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi,
                          double alpha_A, double *E_out) {
    SVL_REQUIRE(c && psi && d_psi, "null argument");
    SVL_TRY(check_state(psi, abei, ab, epsf, true));
    SVL_REQUIRE(d_psi->kind == SVL_NODE_C && (!solveA || (d_A && ab && d_A->kind == SVL_EDGE)), "bad direction buffers");
    if (c->opt_cg_fused)
        return svl_cgf_end(c, solveA, kappa2, eps, epsf, H, psi, abei, ab, d_psi, d_A, alpha_psi, alpha_A, E_out);
    if (c->rsize == 4) SVL_TRY(axy3_t<float>(c, d_psi, psi, psi, solveA ? d_A : nullptr, ab, ab, alpha_psi, alpha_A, nullptr, 1.0));
    else SVL_TRY(axy3_t<double>(c, d_psi, psi, psi, solveA ? d_A : nullptr, ab, ab, alpha_psi, alpha_A, nullptr, 1.0));
    if (E_out) return svl_free_energy(c, kappa2, eps, epsf, H, psi, abei, ab, E_out);
    return 0;
}
