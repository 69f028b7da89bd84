// Observables: induced magnetic field, total current, supercurrent, and the GPU candidate pass
// of the vortex detector.  Reference: svirl/cuda/observables.h:5-235,
// svirl/observables/vortex_detector.py:55-72.
#include "common.cuh"

template <typename R>
__global__ void __launch_bounds__(256)
k_bfield(Geo g, const R *__restrict__ ae, const R *__restrict__ be, const R *__restrict__ a, const R *__restrict__ b,
         R *__restrict__ B) {
    const R idx = (R)g.idx, idy = (R)g.idy;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Nx - 1 || j >= g.j1 || j >= g.Ny - 1) return;
    size_t n = g.at(i, j);
    R v = 0;
    if (ae) v += idx * (be[n + 1] - be[n]) - idy * (ae[n + g.P] - ae[n]);
    if (a) v += idx * (b[n + 1] - b[n]) - idy * (a[n + g.P] - a[n]);
    B[n] = v;
}

// j = kappa2 * (curl curl stencil with boundary doubling): observables.h:66-155 (quirk Q10)
template <typename R>
__global__ void __launch_bounds__(256)
k_current(Geo g, R kappa2, R H, const R *__restrict__ ae, const R *__restrict__ be, const R *__restrict__ a,
          const R *__restrict__ b, R *__restrict__ oa, R *__restrict__ ob) {
    const R idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2, idxy = (R)g.idxy;
    const int P = g.P;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Nx || j >= g.j1) return;
    size_t n = g.at(i, j);
    if (i < g.Nx - 1) {
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
        oa[n] = kappa2 * v;
    }
    if (j < g.Ny - 1) {
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
        ob[n] = kappa2 * v;
    }
}

// js = w_DU * Im(conj(psi0) U(d A) psi1) / d   (observables.h:198-234)
template <typename R>
__global__ void __launch_bounds__(256)
k_supercurrent(Geo g, const uint8_t *__restrict__ nf, const typename V2<R>::type *__restrict__ psi,
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
        R jl = 0;
        if (f & (NF_PM | NF_PP)) {
            R A = 0;
            if (ae) A += ae[n];
            if (a) A += a[n];
            jl = ((R)0.5 * (pm + pp)) * idx * js_link<R, C>(p0, dx * A, psi[n + 1]);
        }
        oa[n] = jl;
    }
    if (j < g.Ny - 1) {
        R jl = 0;
        if (f & (NF_MP | NF_PP)) {
            R A = 0;
            if (be) A += be[n];
            if (b) A += b[n];
            jl = ((R)0.5 * (mp + pp)) * idy * js_link<R, C>(p0, dy * A, psi[n + g.P]);
        }
        ob[n] = jl;
    }
}

// Winding number per cell in double; emits a SUPERSET of the reference's candidates
// (|v| > 0.5 && |v - round v| < 0.1 on the host) using looser thresholds.
template <typename R>
__global__ void __launch_bounds__(256)
k_winding(Geo g, double H, const typename V2<R>::type *__restrict__ psi, const R *__restrict__ a,
          const R *__restrict__ b, long long *cells, double *vals, unsigned long long *count, size_t cap) {
    typedef typename V2<R>::type C;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = g.j0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.Nx - 1 || j >= g.j1 || j >= g.Ny - 1) return;
    size_t n = g.at(i, j);
    const double PI = 3.14159265358979323846, TWO_PI = 2.0 * PI;
    C p00 = psi[n], pp0 = psi[n + 1], ppp = psi[n + g.P + 1], p0p = psi[n + g.P];
    double t00 = atan2((double)p00.y, (double)p00.x), tp0 = atan2((double)pp0.y, (double)pp0.x);
    double tpp = atan2((double)ppp.y, (double)ppp.x), t0p = atan2((double)p0p.y, (double)p0p.x);
    double dx = g.dx, dy = g.dy;
    auto m = [&](double x) { double r = fmod(x, TWO_PI); return r < 0.0 ? r + TWO_PI : r; };
    double v = -(0.5 / PI) * (m(tp0 - t00 - dx * (double)a[n] + PI) + m(tpp - tp0 - dy * (double)b[n + 1] + PI)
                              + m(t0p - tpp + dx * (double)a[n + g.P] + PI) + m(t00 - t0p + dy * (double)b[n] + PI)
                              - 4.0 * PI + dx * dy * H);
    if (fabs(v) > 0.45 && fabs(v - rint(v)) < 0.15) {
        unsigned long long k = atomicAdd(count, 1ull);
        if (k < cap) { cells[k] = (long long)i + (long long)(g.Nx - 1) * j; vals[k] = v; }
    }
}

#define GRID2D(c) dim3 bdim(32, 8), gdim(((c)->g.Nx + 31) / 32, ((c)->g.j1 - (c)->g.j0 + 7) / 8)
#define EDGE_A(buf, R) ((buf) ? (const R *)(buf)->p[0] : nullptr)
#define EDGE_B(buf, R) ((buf) ? (const R *)(buf)->p[1] : nullptr)

extern "C" int svl_magnetic_field(svl_ctx *c, const svl_buf *abei, const svl_buf *ab, svl_buf *B) {
    SVL_REQUIRE(c && B && B->kind == SVL_CELL_R, "B must be SVL_CELL_R");
    SVL_REQUIRE((!abei || abei->kind == SVL_EDGE) && (!ab || ab->kind == SVL_EDGE), "edge buffers required");
    GRID2D(c);
    if (c->rsize == 4) k_bfield<float><<<gdim, bdim, 0, c->stream>>>(c->g, EDGE_A(abei, float), EDGE_B(abei, float), EDGE_A(ab, float), EDGE_B(ab, float), (float *)B->p[0]);
    else k_bfield<double><<<gdim, bdim, 0, c->stream>>>(c->g, EDGE_A(abei, double), EDGE_B(abei, double), EDGE_A(ab, double), EDGE_B(ab, double), (double *)B->p[0]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

extern "C" int svl_current_density(svl_ctx *c, double kappa2, double H, const svl_buf *abei, const svl_buf *ab, svl_buf *out) {
    SVL_REQUIRE(c && out && out->kind == SVL_EDGE, "out must be SVL_EDGE");
    SVL_REQUIRE((!abei || abei->kind == SVL_EDGE) && (!ab || ab->kind == SVL_EDGE), "edge buffers required");
    GRID2D(c);
    if (c->rsize == 4) k_current<float><<<gdim, bdim, 0, c->stream>>>(c->g, (float)kappa2, (float)H, EDGE_A(abei, float), EDGE_B(abei, float), EDGE_A(ab, float), EDGE_B(ab, float), (float *)out->p[0], (float *)out->p[1]);
    else k_current<double><<<gdim, bdim, 0, c->stream>>>(c->g, kappa2, H, EDGE_A(abei, double), EDGE_B(abei, double), EDGE_A(ab, double), EDGE_B(ab, double), (double *)out->p[0], (double *)out->p[1]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

extern "C" int svl_supercurrent_density(svl_ctx *c, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *out) {
    SVL_REQUIRE(c && psi && psi->kind == SVL_NODE_C && out && out->kind == SVL_EDGE, "psi NODE_C and out EDGE required");
    SVL_REQUIRE((!abei || abei->kind == SVL_EDGE) && (!ab || ab->kind == SVL_EDGE), "edge buffers required");
    GRID2D(c);
    if (c->rsize == 4) k_supercurrent<float><<<gdim, bdim, 0, c->stream>>>(c->g, c->nf, (const float2 *)psi->p[0], EDGE_A(abei, float), EDGE_B(abei, float), EDGE_A(ab, float), EDGE_B(ab, float), (float *)out->p[0], (float *)out->p[1]);
    else k_supercurrent<double><<<gdim, bdim, 0, c->stream>>>(c->g, c->nf, (const double2 *)psi->p[0], EDGE_A(abei, double), EDGE_B(abei, double), EDGE_A(ab, double), EDGE_B(ab, double), (double *)out->p[0], (double *)out->p[1]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

extern "C" int svl_vortex_candidates(svl_ctx *c, double H, const svl_buf *psi, const svl_buf *ab, int64_t *cells_out,
                                     double *v_out, size_t max_out, size_t *count_out) {
    SVL_REQUIRE(c && psi && ab && count_out, "null argument");
    SVL_REQUIRE(psi->kind == SVL_NODE_C && ab->kind == SVL_EDGE, "psi NODE_C and ab EDGE required");
    if (c->cand_cap < max_out || !c->d_cand) {
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_cand); cudaFree(c->d_candv);
        c->cand_cap = max_out ? max_out : 1;
        SVL_CHECK(cudaMalloc(&c->d_cand, c->cand_cap * sizeof(long long)));
        SVL_CHECK(cudaMalloc(&c->d_candv, c->cand_cap * sizeof(double)));
    }
    SVL_CHECK(cudaMemsetAsync(c->d_ncand, 0, sizeof(unsigned long long), c->stream));
    GRID2D(c);
    if (c->rsize == 4) k_winding<float><<<gdim, bdim, 0, c->stream>>>(c->g, H, (const float2 *)psi->p[0], (const float *)ab->p[0], (const float *)ab->p[1], c->d_cand, c->d_candv, c->d_ncand, max_out);
    else k_winding<double><<<gdim, bdim, 0, c->stream>>>(c->g, H, (const double2 *)psi->p[0], (const double *)ab->p[0], (const double *)ab->p[1], c->d_cand, c->d_candv, c->d_ncand, max_out);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    unsigned long long cnt = 0;
    SVL_CHECK(cudaMemcpyAsync(&cnt, c->d_ncand, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    *count_out = (size_t)cnt;
    size_t m = cnt < max_out ? (size_t)cnt : max_out;
    if (m && cells_out) SVL_CHECK(cudaMemcpy(cells_out, c->d_cand, m * sizeof(long long), cudaMemcpyDeviceToHost));
    if (m && v_out) SVL_CHECK(cudaMemcpy(v_out, c->d_candv, m * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
