// Small grids: ALL time steps of a td() call in ONE launch of one thread-block cluster.
//
// At the README size (129^2, BASELINE configs[0]) a Jacobi sweep is a microsecond of work and the step time is launch
// and synchronisation latency: the reference pays fill + launch + blocking read-back per sweep (svirl/solvers/td.py:164-202,
// 274-311: ~1 ms per step), the batched drivers of td.cu two host round trips per step (0.105 ms).  Here a cluster of up
// to 16 CTAs (hardware cluster barrier, ~0.3 us) keeps the whole solve on the device: every thread owns up to NPT nodes
// with their constants in registers (right-hand side, the four link coefficients w*dt/d^2*exp(-+i d A) -- one sincos per
// link and SOLVE --, 1/diagonal); per sweep it reads the four neighbours from L2 (ld.cg: the iterate is written by other
// SMs), writes its nodes, the max-norm update goes through one atomicMax per CTA, and after the cluster barrier every
// thread evaluates the reference's stop test (td.h:124-132 + td.py:198-201) on the same word.  psi-solve, A-solve (with the
// link-phase aliasing quirk Q1), Langevin noise and the rand_t bookkeeping follow td.cu / the reference exactly; the
// sweep counts are the reference's (tests: README fixture, 1000 steps).  The host synchronises once per td() call.
// This is the "graphs" option of the library (north_star: launch overhead on small grids), default on; grids above
// 32768 nodes, row slabs and fixed vortices take the batched drivers.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define TS_THREADS 512
#define TS_MAX_CTAS 16
#define TS_MAX_NPT 4

template <typename R> struct SmallArgs {
    Geo g;
    int Nt, solveA;
    R dt, eps, kappa2, rho, H, lang_psi, lang_A;
    double stop_psi, stop_A;
    const R *epsf;
    const uint8_t *nf;
    typename V2<R>::type *psi[3];        // [0] the caller's buffer (state), [1], [2] scratch
    R *a[3], *b[3];
    uint32_t rand_t;
    unsigned long long *ring;            // 4 residual words, zero on entry
    long long *out;                      // psi sweeps, A sweeps, index of the buffer holding psi / A at the end, rand_t
};

__device__ __forceinline__ bool ts_stop(double r, double eps, bool fp32) {
    // exact reference decision, as td.cu:stop_rule
    double v = fp32 ? (double)(float)(1.0e4 * r / (double)(float)eps) : 1.0e4 * r / eps;
    if (v > 1.0e8) v = 1.0e8;
    return (int)v < 10000;
}
template <typename T> __device__ __forceinline__ T ts_ld(const T *p) { return __ldcg(p); }

template <typename R, int NPT>
__global__ void __launch_bounds__(TS_THREADS, 1) k_td_small(const __grid_constant__ SmallArgs<R> A) {
    typedef typename V2<R>::type C;
    cg::cluster_group cl = cg::this_cluster();
    const Geo &g = A.g;
    const int T = gridDim.x * TS_THREADS, gt = blockIdx.x * TS_THREADS + threadIdx.x;
    const int P = g.P;
    const bool fp32 = sizeof(R) == 4;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2, idxy = (R)g.idxy;
    const R dt = A.dt, cx = dt * idx2, cy = dt * idy2;

    int off[NPT], ci[NPT], cj[NPT];
    unsigned fl[NPT];
    bool ok[NPT];
#pragma unroll
    for (int k = 0; k < NPT; k++) {
        const int nl = gt + k * T;
        ok[k] = nl < g.Nx * g.Ny;
        ci[k] = ok[k] ? nl % g.Nx : 0;
        cj[k] = ok[k] ? nl / g.Nx : 0;
        off[k] = (cj[k] - g.rb) * P + ci[k];
        fl[k] = ok[k] ? A.nf[off[k]] : 0u;
    }
    unsigned sidx = 0;                               // sweeps done by this launch: index into the residual ring
    int pr[3] = {0, 1, 2}, ar[3] = {0, 1, 2};        // buffer roles: [0] state / iterate 0, [1], [2] ping-pong
    long long npsi = 0, nA = 0;
    uint32_t rand_t = A.rand_t;

    // one sweep's epilogue: CTA max -> ring word, cluster barrier, everybody reads the same maximum
    auto finish_sweep = [&](double rmax) {
        block_max_to_slot(rmax, A.ring + (sidx & 3));
        if (gt == 0) A.ring[(sidx + 2) & 3] = 0ull;  // free since everybody passed the previous barrier
        cl.sync();
        const unsigned long long bits = ts_ld(A.ring + (sidx & 3));
        sidx++;
        return __longlong_as_double((long long)bits);
    };

    for (int step = 0; step < A.Nt; step++) {
        // ============================ psi solve (td.h:5-133; drivers td.py:157-218)
        {
            C *B0 = A.psi[pr[0]], *S1 = A.psi[pr[1]], *S2 = A.psi[pr[2]];
            const R *pa = A.a[ar[0]], *pb = A.b[ar[0]];
            C q[NPT], own[NPT], LW[NPT], LE[NPT], LS[NPT], LN[NPT];
            R di[NPT];
            const bool noise = A.lang_psi > (R)1.0e-32;
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                C z; z.x = 0; z.y = 0;
                q[k] = z; own[k] = z; LW[k] = z; LE[k] = z; LS[k] = z; LN[k] = z; di[k] = 0;
                if (!ok[k]) continue;
                const int n = off[k];
                const unsigned f = fl[k];
                own[k] = ts_ld(B0 + n);
                if (!f) continue;
                C qq = own[k];
                if (noise) {
                    const uint32_t nn = (uint32_t)ci[k] + (uint32_t)g.Nx * (uint32_t)cj[k];
                    qq.x += A.lang_psi * (rand_1<R>(nn, rand_t) - (R)0.5);
                    qq.y += A.lang_psi * (rand_2<R>(nn, rand_t) - (R)0.5);
                }
                const bool wW = f & (NF_MM | NF_MP), wE = f & (NF_PM | NF_PP), wS = f & (NF_MM | NF_PM), wN = f & (NF_MP | NF_PP);
                R sn, cs;
                if (wW) { sincos_r<R>(dx * ts_ld(pa + n - 1), &sn, &cs); LW[k].x = cx * cs; LW[k].y = cx * sn; }
                if (wE) { sincos_r<R>(dx * ts_ld(pa + n), &sn, &cs); LE[k].x = cx * cs; LE[k].y = cx * sn; }
                if (wS) { sincos_r<R>(dy * ts_ld(pb + n - P), &sn, &cs); LS[k].x = cy * cs; LS[k].y = cy * sn; }
                if (wN) { sincos_r<R>(dy * ts_ld(pb + n), &sn, &cs); LN[k].x = cy * cs; LN[k].y = cy * sn; }
                const R e = A.epsf ? A.epsf[n] : A.eps;
                const R nwx = (wW ? (R)1 : (R)0) + (wE ? (R)1 : (R)0), nwy = (wS ? (R)1 : (R)0) + (wN ? (R)1 : (R)0);
                const R D = (R)1.0 + dt * (qq.x * qq.x + qq.y * qq.y - e + (idx2 * nwx + idy2 * nwy));
                di[k] = rcp_r(D);
                q[k] = qq;
            }
            int res = SVL_MAX_SWEEPS;
            for (int s = 0; s < SVL_MAX_SWEEPS; s++) {
                const C *in = s == 0 ? B0 : ((s & 1) ? S1 : S2);
                C *out = (s & 1) ? S2 : S1;
                double rmax = 0.0;
#pragma unroll
                for (int k = 0; k < NPT; k++) {
                    if (!ok[k]) continue;
                    const int n = off[k];
                    const C pw = ts_ld(in + n - 1), pe = ts_ld(in + n + 1), pS = ts_ld(in + n - P), pN = ts_ld(in + n + P);
                    // W,S use (c + i s) psi, E,N use (c - i s) psi (the order of psi_tile.cu)
                    R ax = q[k].x, ay = q[k].y;
                    ax = fma_r(LW[k].x, pw.x, ax);  ay = fma_r(LW[k].x, pw.y, ay);
                    ax = fma_r(-LW[k].y, pw.y, ax); ay = fma_r(LW[k].y, pw.x, ay);
                    ax = fma_r(LE[k].x, pe.x, ax);  ay = fma_r(LE[k].x, pe.y, ay);
                    ax = fma_r(LE[k].y, pe.y, ax);  ay = fma_r(-LE[k].y, pe.x, ay);
                    ax = fma_r(LS[k].x, pS.x, ax);  ay = fma_r(LS[k].x, pS.y, ay);
                    ax = fma_r(-LS[k].y, pS.y, ax); ay = fma_r(LS[k].y, pS.x, ay);
                    ax = fma_r(LN[k].x, pN.x, ax);  ay = fma_r(LN[k].x, pN.y, ay);
                    ax = fma_r(LN[k].y, pN.y, ax);  ay = fma_r(-LN[k].y, pN.x, ay);
                    C nx;
                    nx.x = ax * di[k]; nx.y = ay * di[k];
                    out[n] = nx;
                    rmax = fmax(rmax, (double)fmax(fabs(nx.x - own[k].x), fabs(nx.y - own[k].y)));
                    own[k] = nx;
                }
                const double r = finish_sweep(rmax);
                if (ts_stop(r, A.stop_psi, fp32)) { res = s + 1; break; }
            }
            npsi += res;
            const int w = ((res - 1) & 1) ? 2 : 1;           // role that holds the result
            const int t0 = pr[0]; pr[0] = pr[w]; pr[w] = t0;
            rand_t += 1u;                                    // td.py:204
        }
        // ============================ A solve (td.h:311-463; drivers td.py:252-325; quirk Q1)
        if (A.solveA) {
            const C *psi = A.psi[pr[0]];
            R *B0a = A.a[ar[0]], *B0b = A.b[ar[0]], *S1a = A.a[ar[1]], *S1b = A.b[ar[1]], *S2a = A.a[ar[2]], *S2b = A.b[ar[2]];
            const R dt_rho = dt * A.rho, dtrk = dt_rho * A.kappa2;
            const R inv_da = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idy2), inv_db = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idx2);
            const bool noise = A.lang_A > (R)1.0e-32;
            C p0[NPT], pE[NPT], pN[NPT];
            R qa[NPT], qb[NPT], owa[NPT], owb[NPT], ca[NPT], cb[NPT];
            // boundary terms of an edge (td.h:375-377, 421-423): recomputed where needed, two compares each
            auto bnd_a = [&](int j, R &rh, R &dd) {
                rh = 0; dd = 1;
                if (j == 0) { rh = (R)2.0 * A.kappa2 * A.H * idy; dd = 2; }
                else if (j + 1 == g.Ny) { rh = -(R)2.0 * A.kappa2 * A.H * idy; dd = 2; }
            };
            auto bnd_b = [&](int i, R &rh, R &dd) {
                rh = 0; dd = 1;
                if (i == 0) { rh = -(R)2.0 * A.kappa2 * A.H * idx; dd = 2; }
                else if (i + 1 == g.Nx) { rh = (R)2.0 * A.kappa2 * A.H * idx; dd = 2; }
            };
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                C z; z.x = 0; z.y = 0;
                p0[k] = z; pE[k] = z; pN[k] = z;
                qa[k] = 0; qb[k] = 0; owa[k] = 0; owb[k] = 0; ca[k] = 0; cb[k] = 0;
                if (!ok[k]) continue;
                const int n = off[k], i = ci[k], j = cj[k];
                p0[k] = ts_ld(psi + n); pE[k] = ts_ld(psi + n + 1); pN[k] = ts_ld(psi + n + P);
                if (i < g.Nx - 1) {
                    owa[k] = ts_ld(B0a + n);
                    qa[k] = owa[k];
                    if (noise) qa[k] += A.lang_A * (rand_1<R>((uint32_t)i + (uint32_t)(g.Nx - 1) * (uint32_t)j, rand_t) - (R)0.5);
                }
                if (j < g.Ny - 1) {
                    owb[k] = ts_ld(B0b + n);
                    qb[k] = owb[k];
                    if (noise)
                        qb[k] += A.lang_A * (rand_2<R>((uint32_t)((size_t)(g.Nx - 1) * g.Ny) + (uint32_t)i + (uint32_t)g.Nx * (uint32_t)j, rand_t) - (R)0.5);
                }
            }
            int res = SVL_MAX_SWEEPS;
            for (int s = 0; s < SVL_MAX_SWEEPS; s++) {
                const R *ina = s == 0 ? B0a : ((s & 1) ? S1a : S2a), *inb = s == 0 ? B0b : ((s & 1) ? S1b : S2b);
                R *oa = (s & 1) ? S2a : S1a, *ob = (s & 1) ? S2b : S1b;
                double rmax = 0.0;
#pragma unroll
                for (int k = 0; k < NPT; k++) {
                    if (!ok[k]) continue;
                    const int n = off[k], i = ci[k], j = cj[k];
                    const unsigned f = fl[k];
                    R rha, dda, rhb, ddb;
                    bnd_a(j, rha, dda);
                    bnd_b(i, rhb, ddb);
                    if (i < g.Nx - 1) {
                        if (!(s & 1)) {      // the link phase of sweeps 2m and 2m+1 is iterate 2m (quirk Q1)
                            R jl = 0;
                            if (f & (NF_PM | NF_PP)) jl = idx * js_link<R, C>(p0[k], dx * owa[k], pE[k]);
                            ca[k] = qa[k] + dt_rho * (jl + rha);
                        }
                        R lo = 0, hi = 0;
                        if (j > 0) lo = idy2 * ts_ld(ina + n - P) - idxy * ts_ld(inb + n - P) + idxy * ts_ld(inb + n - P + 1);
                        if (j + 1 < g.Ny) hi = idy2 * ts_ld(ina + n + P) + idxy * owb[k] - idxy * ts_ld(inb + n + 1);
                        const R nx = (ca[k] + dtrk * dda * (lo + hi)) * inv_da;
                        oa[n] = nx;
                        rmax = fmax(rmax, fabs((double)(nx - owa[k])));
                        // own a-value of the INPUT iterate is still needed by the b-edge below: keep it until then
                        const R olda = owa[k];
                        owa[k] = nx;
                        if (j < g.Ny - 1) {
                            if (!(s & 1)) {
                                R jl = 0;
                                if (f & (NF_MP | NF_PP)) jl = idy * js_link<R, C>(p0[k], dy * owb[k], pN[k]);
                                cb[k] = qb[k] + dt_rho * (jl + rhb);
                            }
                            R lo2 = 0, hi2 = 0;
                            if (i > 0) lo2 = idx2 * ts_ld(inb + n - 1) - idxy * ts_ld(ina + n - 1) + idxy * ts_ld(ina + n - 1 + P);
                            if (i + 1 < g.Nx) hi2 = idx2 * ts_ld(inb + n + 1) + idxy * olda - idxy * ts_ld(ina + n + P);
                            const R nb = (cb[k] + dtrk * ddb * (lo2 + hi2)) * inv_db;
                            ob[n] = nb;
                            rmax = fmax(rmax, fabs((double)(nb - owb[k])));
                            owb[k] = nb;
                        }
                    } else if (j < g.Ny - 1) {       // last column: only the b-edge exists
                        if (!(s & 1)) {
                            R jl = 0;
                            if (f & (NF_MP | NF_PP)) jl = idy * js_link<R, C>(p0[k], dy * owb[k], pN[k]);
                            cb[k] = qb[k] + dt_rho * (jl + rhb);
                        }
                        R lo2 = 0;
                        if (i > 0) lo2 = idx2 * ts_ld(inb + n - 1) - idxy * ts_ld(ina + n - 1) + idxy * ts_ld(ina + n - 1 + P);
                        const R nb = (cb[k] + dtrk * ddb * (lo2 + (R)0)) * inv_db;
                        ob[n] = nb;
                        rmax = fmax(rmax, fabs((double)(nb - owb[k])));
                        owb[k] = nb;
                    }
                }
                const double r = finish_sweep(rmax);
                if (ts_stop(r, A.stop_A, fp32)) { res = s + 1; break; }
            }
            nA += res;
            const int w = ((res - 1) & 1) ? 2 : 1;
            const int t0 = ar[0]; ar[0] = ar[w]; ar[w] = t0;
            rand_t += 1u;                                    // td.py:313
        }
    }
    if (gt == 0) {
        A.out[0] = npsi; A.out[1] = nA; A.out[2] = pr[0]; A.out[3] = ar[0]; A.out[4] = (long long)rand_t;
    }
}

template <typename R, int NPT>
static int ts_launch(svl_ctx *c, const SmallArgs<R> &A, int nctas, bool *handled) {
    auto kern = k_td_small<R, NPT>;
    if (nctas > 8) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(nctas); lc.blockDim = dim3(TS_THREADS); lc.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kern, &lc) != cudaSuccess || ncl < 1) { cudaGetLastError(); return 0; }
    SVL_CHECK(cudaLaunchKernelEx(&lc, kern, A));
    *handled = true;
    return 0;
}

template <typename R>
static int ts_run_t(svl_ctx *c, int Nt, double dt, int solveA, double eps, const svl_buf *epsf, double kappa2, double rho,
                    double H, svl_buf *psi, svl_buf *ab, double lang_psi, double lang_A, uint32_t *rand_t, double stop_psi,
                    double stop_A, long long *sweeps, bool *handled) {
    typedef typename V2<R>::type C;
    const Geo &g = c->g;
    const long N = (long)g.Nx * g.Ny;
    int npt = 0, nctas = 0;
    for (int k = 1; k <= TS_MAX_NPT && !npt; k++)
        for (int m = 1; m <= TS_MAX_CTAS; m *= 2)
            if ((long)m * TS_THREADS * k >= N) { npt = k; nctas = m; break; }
    if (!npt) return 0;
    svl_buf *ps[2], *as[2];
    for (int k = 0; k < 2; k++) { SVL_TRY(svl_scratch_node(c, k, &ps[k])); SVL_TRY(svl_scratch_edge(c, k, &as[k])); }
    SmallArgs<R> A;
    memset(&A, 0, sizeof(A));
    A.g = g; A.Nt = Nt; A.solveA = solveA;
    A.dt = (R)dt; A.eps = (R)eps; A.kappa2 = (R)kappa2; A.rho = (R)rho; A.H = (R)H; A.lang_psi = (R)lang_psi; A.lang_A = (R)lang_A;
    A.stop_psi = stop_psi; A.stop_A = stop_A;
    A.epsf = epsf ? (const R *)epsf->p[0] : nullptr;
    A.nf = c->nf;
    svl_buf *pb[3] = {psi, ps[0], ps[1]}, *abb[3] = {ab, as[0], as[1]};
    for (int k = 0; k < 3; k++) { A.psi[k] = (C *)pb[k]->p[0]; A.a[k] = (R *)abb[k]->p[0]; A.b[k] = (R *)abb[k]->p[1]; }
    A.rand_t = *rand_t;
    A.ring = c->d_resid;
    long long *dout = (long long *)(c->d_result + 48);
    A.out = dout;
    SVL_CHECK(cudaMemsetAsync(c->d_resid, 0, 4 * sizeof(unsigned long long), c->stream));
    switch (npt) {
        case 1: SVL_TRY((ts_launch<R, 1>(c, A, nctas, handled))); break;
        case 2: SVL_TRY((ts_launch<R, 2>(c, A, nctas, handled))); break;
        case 3: SVL_TRY((ts_launch<R, 3>(c, A, nctas, handled))); break;
        default: SVL_TRY((ts_launch<R, 4>(c, A, nctas, handled))); break;
    }
    if (!*handled) return 0;
    c->stat_launches += 1;
    long long *hout = (long long *)(c->h_result + 48);
    SVL_CHECK(cudaMemcpyAsync(hout, dout, 5 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    if (hout[2] != 0) SVL_TRY(svl_swap(c, psi, pb[hout[2]]));
    if (hout[3] != 0) SVL_TRY(svl_swap(c, ab, abb[hout[3]]));
    *rand_t = (uint32_t)hout[4];
    if (sweeps) { sweeps[0] += hout[0]; sweeps[1] += hout[1]; }
    c->stat_psi_sweeps += (double)hout[0]; c->stat_A_sweeps += (double)hout[1];
    if (Nt > 0) { c->pred_psi2 = c->pred_psi = (int)(hout[0] / Nt); c->pred_A2 = c->pred_A = (int)(hout[1] / Nt); }
    return 0;
}

// Called by svl_td_run: runs the whole call on the device when the grid is small enough; *handled = false otherwise.
int svl_td_small_run(svl_ctx *c, int Nt, double dt, int solveA, double eps, const svl_buf *epsf, double kappa2, double rho,
                     double H, svl_buf *psi, svl_buf *ab, double lang_psi, double lang_A, uint32_t *rand_t, double stop_psi,
                     double stop_A, long long *sweeps, bool *handled) {
    *handled = false;
    if (!c->opt_graphs || c->slab_on || Nt <= 0) return 0;
    if ((long)c->g.Nx * c->g.Ny > (long)TS_MAX_CTAS * TS_THREADS * TS_MAX_NPT) return 0;
    if (c->rsize == 4) return ts_run_t<float>(c, Nt, dt, solveA, eps, epsf, kappa2, rho, H, psi, ab, lang_psi, lang_A, rand_t,
                                              stop_psi, stop_A, sweeps, handled);
    return ts_run_t<double>(c, Nt, dt, solveA, eps, epsf, kappa2, rho, H, psi, ab, lang_psi, lang_A, rand_t, stop_psi, stop_A,
                            sweeps, handled);
}
