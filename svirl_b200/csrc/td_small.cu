// Small grids: ALL time steps of a td() call in ONE launch of one thread-block cluster.
//
// At the README size (129^2, BASELINE configs[0]) a Jacobi sweep is a microsecond of work and the step time is launch
// and synchronisation latency: the reference pays fill + launch + blocking read-back per sweep (svirl/solvers/td.py:164-202,
// 274-311: ~1 ms per step), the batched drivers of td.cu two host round trips per step (0.105 ms).  Here a cluster of up
// to 16 CTAs (hardware cluster barrier) keeps the whole solve on the device: every thread owns up to NPT nodes with their
// constants in registers (right-hand side, the four link coefficients w*dt/d^2*exp(-+i d A) -- one sincos per link and
// SOLVE --, 1/diagonal); the iterate lives in shared memory (halo rows pushed to the neighbouring CTAs through
// distributed shared memory), the per-CTA max-norm updates are broadcast the same way, and after the cluster barrier
// every thread evaluates the reference's stop test (td.h:124-132 + td.py:198-201) on the same numbers.  psi-solve, A-solve (with the
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
    int M;                               // nodes per CTA (linear node index n = rank * M + m, x fastest); M >= Nx
    R dt, eps, kappa2, rho, H, lang_psi, lang_A;
    double stop_psi, stop_A;
    const R *epsf;
    const uint8_t *nf;
    typename V2<R>::type *psi;           // state, updated in place
    R *a, *b;
    uint32_t rand_t;
    long long *out;                      // psi sweeps, A sweeps, rand_t
};

__device__ __forceinline__ bool ts_stop(double r, double eps, bool fp32) {
    // exact reference decision, as td.cu:stop_rule
    double v = fp32 ? (double)(float)(1.0e4 * r / (double)(float)eps) : 1.0e4 * r / eps;
    if (v > 1.0e8) v = 1.0e8;
    return (int)v < 10000;
}

// The iterate lives in SHARED memory: CTA `rank` owns the M consecutive nodes [rank*M, rank*M + M) and keeps them, plus a
// halo of Nx nodes on either side (the rows above and below), in a double-buffered array; a sweep reads only local shared
// memory, writes its nodes locally and -- for the first / last Nx nodes -- into the neighbouring CTA's halo through
// distributed shared memory.  Global memory is touched at the start and at the end of a solve.
template <typename R, int NPT>
__global__ void __launch_bounds__(TS_THREADS, 1) k_td_small(const __grid_constant__ SmallArgs<R> A) {
    typedef typename V2<R>::type C;
    cg::cluster_group cl = cg::this_cluster();
    const Geo &g = A.g;
    const int tid = threadIdx.x, rank = (int)cl.block_rank(), nct = (int)cl.num_blocks();
    const int Nx = g.Nx, P = g.P, M = A.M, N = g.Nx * g.Ny, W = M + 2 * Nx;
    const bool fp32 = sizeof(R) == 4;
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2, idxy = (R)g.idxy;
    const R dt = A.dt, cx = dt * idx2, cy = dt * idy2;
    extern __shared__ __align__(16) unsigned char ts_smem[];
    C *X0 = (C *)ts_smem, *X1 = X0 + W;
    unsigned long long *red = (unsigned long long *)(X1 + W);          // [2][TS_MAX_CTAS] per-CTA maxima of a sweep
    // the neighbours' copies (halo pushes) and everybody's reduction slots
    C *lo0 = rank > 0 ? cl.map_shared_rank(X0, rank - 1) : nullptr, *lo1 = rank > 0 ? cl.map_shared_rank(X1, rank - 1) : nullptr;
    C *hi0 = rank + 1 < nct ? cl.map_shared_rank(X0, rank + 1) : nullptr, *hi1 = rank + 1 < nct ? cl.map_shared_rank(X1, rank + 1) : nullptr;

    int off[NPT], ci[NPT], cj[NPT], L[NPT];
    unsigned fl[NPT];
    bool ok[NPT];
#pragma unroll
    for (int k = 0; k < NPT; k++) {
        const int m = tid + k * TS_THREADS, n = rank * M + m;
        ok[k] = m < M && n < N;
        ci[k] = ok[k] ? n % Nx : 0;
        cj[k] = ok[k] ? n / Nx : 0;
        off[k] = (cj[k] - g.rb) * P + ci[k];
        L[k] = m + Nx;
        fl[k] = ok[k] ? A.nf[off[k]] : 0u;
    }
    unsigned sidx = 0;
    long long npsi = 0, nA = 0;
    uint32_t rand_t = A.rand_t;

    // fill X0 (own nodes + both halos) from a global plane through `get`, clear X1; ends with a cluster barrier so that
    // nobody pushes into a buffer that is still being initialised
    auto load_state = [&](auto get) {
        for (int q = tid; q < W; q += TS_THREADS) {
            const int n = rank * M - Nx + q;
            C v; v.x = 0; v.y = 0;
            if (n >= 0 && n < N) v = get((n / Nx - g.rb) * P + n % Nx);
            X0[q] = v;
            C z; z.x = 0; z.y = 0;
            X1[q] = z;
        }
        cl.sync();
    };
    // store one node of the new iterate: locally and into the neighbour's halo
    auto put = [&](C *out, C *olo, C *ohi, int k, C v) {
        const int m = L[k] - Nx;
        out[L[k]] = v;
        if (olo && m < Nx) olo[M + Nx + m] = v;                 // my first Nx nodes = upper halo of rank-1
        if (ohi && m >= M - Nx) ohi[m - (M - Nx)] = v;          // my last Nx nodes = lower halo of rank+1
    };
    // one sweep's epilogue: CTA max -> everybody's slot, cluster barrier, maximum over the CTAs
    auto finish_sweep = [&](double rmax) {
        __shared__ double sm_w[TS_THREADS / 32];
        rmax = warp_max(rmax);
        if ((tid & 31) == 0) sm_w[tid >> 5] = rmax;
        __syncthreads();
        unsigned long long *slot = red + (sidx & 1) * TS_MAX_CTAS;
        if (tid < 32) {
            double r = tid < TS_THREADS / 32 ? sm_w[tid] : 0.0;
            r = warp_max(r);
            r = __shfl_sync(0xffffffffu, r, 0);
            if (tid < nct) cl.map_shared_rank(slot, tid)[rank] = (unsigned long long)__double_as_longlong(r);
        }
        cl.sync();
        double r = 0.0;
        for (int q = 0; q < nct; q++) r = fmax(r, __longlong_as_double((long long)slot[q]));
        sidx++;
        return r;
    };

    for (int step = 0; step < A.Nt; step++) {
        // ============================ psi solve (td.h:5-133; drivers td.py:157-218)
        {
            load_state([&](int o) { return __ldcg(A.psi + o); });
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
                own[k] = X0[L[k]];
                if (!f) continue;
                C qq = own[k];
                if (noise) {
                    const uint32_t nn = (uint32_t)ci[k] + (uint32_t)Nx * (uint32_t)cj[k];
                    qq.x += A.lang_psi * (rand_1<R>(nn, rand_t) - (R)0.5);
                    qq.y += A.lang_psi * (rand_2<R>(nn, rand_t) - (R)0.5);
                }
                const bool wW = f & (NF_MM | NF_MP), wE = f & (NF_PM | NF_PP), wS = f & (NF_MM | NF_PM), wN = f & (NF_MP | NF_PP);
                R sn, cs;
                if (wW) { sincos_r<R>(dx * __ldcg(A.a + n - 1), &sn, &cs); LW[k].x = cx * cs; LW[k].y = cx * sn; }
                if (wE) { sincos_r<R>(dx * __ldcg(A.a + n), &sn, &cs); LE[k].x = cx * cs; LE[k].y = cx * sn; }
                if (wS) { sincos_r<R>(dy * __ldcg(A.b + n - P), &sn, &cs); LS[k].x = cy * cs; LS[k].y = cy * sn; }
                if (wN) { sincos_r<R>(dy * __ldcg(A.b + n), &sn, &cs); LN[k].x = cy * cs; LN[k].y = cy * sn; }
                const R e = A.epsf ? A.epsf[n] : A.eps;
                const R nwx = (wW ? (R)1 : (R)0) + (wE ? (R)1 : (R)0), nwy = (wS ? (R)1 : (R)0) + (wN ? (R)1 : (R)0);
                const R D = (R)1.0 + dt * (qq.x * qq.x + qq.y * qq.y - e + (idx2 * nwx + idy2 * nwy));
                di[k] = rcp_r(D);
                q[k] = qq;
            }
            int res = SVL_MAX_SWEEPS;
            for (int s = 0; s < SVL_MAX_SWEEPS; s++) {
                const C *in = (s & 1) ? X1 : X0;
                C *out = (s & 1) ? X0 : X1, *olo = (s & 1) ? lo0 : lo1, *ohi = (s & 1) ? hi0 : hi1;
                double rmax = 0.0;
#pragma unroll
                for (int k = 0; k < NPT; k++) {
                    if (!ok[k]) continue;
                    const int l = L[k];
                    const C pw = in[l - 1], pe = in[l + 1], pS = in[l - Nx], pN = in[l + Nx];
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
                    put(out, olo, ohi, k, nx);
                    rmax = fmax(rmax, (double)fmax(fabs(nx.x - own[k].x), fabs(nx.y - own[k].y)));
                    own[k] = nx;
                }
                const double r = finish_sweep(rmax);
                if (ts_stop(r, A.stop_psi, fp32)) { res = s + 1; break; }
            }
            npsi += res;
#pragma unroll
            for (int k = 0; k < NPT; k++)
                if (ok[k]) A.psi[off[k]] = own[k];
            rand_t += 1u;                                    // td.py:204
            cl.sync();                                       // the new psi is visible to the whole cluster
        }
        // ============================ A solve (td.h:311-463; drivers td.py:252-325; quirk Q1)
        if (A.solveA) {
            // the (a, b) pair of a node travels as one complex-sized element
            load_state([&](int o) { C v; v.x = __ldcg(A.a + o); v.y = __ldcg(A.b + o); return v; });
            const R dt_rho = dt * A.rho, dtrk = dt_rho * A.kappa2;
            const R inv_da = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idy2), inv_db = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idx2);
            const bool noise = A.lang_A > (R)1.0e-32;
            C p0[NPT], pE[NPT], pN[NPT], qab[NPT], own[NPT], cab[NPT];
            // boundary terms of an edge (td.h:375-377, 421-423): recomputed where needed, two compares each
            auto bnd_a = [&](int j, R &rh, R &dd) {
                rh = 0; dd = 1;
                if (j == 0) { rh = (R)2.0 * A.kappa2 * A.H * idy; dd = 2; }
                else if (j + 1 == g.Ny) { rh = -(R)2.0 * A.kappa2 * A.H * idy; dd = 2; }
            };
            auto bnd_b = [&](int i, R &rh, R &dd) {
                rh = 0; dd = 1;
                if (i == 0) { rh = -(R)2.0 * A.kappa2 * A.H * idx; dd = 2; }
                else if (i + 1 == Nx) { rh = (R)2.0 * A.kappa2 * A.H * idx; dd = 2; }
            };
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                C z; z.x = 0; z.y = 0;
                p0[k] = z; pE[k] = z; pN[k] = z; qab[k] = z; own[k] = z; cab[k] = z;
                if (!ok[k]) continue;
                const int n = off[k], i = ci[k], j = cj[k];
                p0[k] = __ldcg(A.psi + n); pE[k] = __ldcg(A.psi + n + 1); pN[k] = __ldcg(A.psi + n + P);
                own[k] = X0[L[k]];
                if (i >= Nx - 1) own[k].x = 0;               // no a-edge in the last column, no b-edge in the last row
                if (j >= g.Ny - 1) own[k].y = 0;
                qab[k] = own[k];
                if (noise && i < Nx - 1) qab[k].x += A.lang_A * (rand_1<R>((uint32_t)i + (uint32_t)(Nx - 1) * (uint32_t)j, rand_t) - (R)0.5);
                if (noise && j < g.Ny - 1)
                    qab[k].y += A.lang_A * (rand_2<R>((uint32_t)((size_t)(Nx - 1) * g.Ny) + (uint32_t)i + (uint32_t)Nx * (uint32_t)j, rand_t) - (R)0.5);
            }
            int res = SVL_MAX_SWEEPS;
            for (int s = 0; s < SVL_MAX_SWEEPS; s++) {
                const C *in = (s & 1) ? X1 : X0;
                C *out = (s & 1) ? X0 : X1, *olo = (s & 1) ? lo0 : lo1, *ohi = (s & 1) ? hi0 : hi1;
                double rmax = 0.0;
#pragma unroll
                for (int k = 0; k < NPT; k++) {
                    if (!ok[k]) continue;
                    const int l = L[k], i = ci[k], j = cj[k];
                    const unsigned f = fl[k];
                    C nx = own[k];
                    if (!(s & 1)) {          // the link phase of sweeps 2m and 2m+1 is iterate 2m (quirk Q1)
                        R rha, dda, rhb, ddb, jl = 0;
                        bnd_a(j, rha, dda);
                        bnd_b(i, rhb, ddb);
                        if (f & (NF_PM | NF_PP)) jl = idx * js_link<R, C>(p0[k], dx * own[k].x, pE[k]);
                        cab[k].x = qab[k].x + dt_rho * (jl + rha);
                        jl = 0;
                        if (f & (NF_MP | NF_PP)) jl = idy * js_link<R, C>(p0[k], dy * own[k].y, pN[k]);
                        cab[k].y = qab[k].y + dt_rho * (jl + rhb);
                    }
                    if (i < Nx - 1) {
                        R rh, dd, lo = 0, hi = 0;
                        bnd_a(j, rh, dd);
                        if (j > 0) lo = idy2 * in[l - Nx].x - idxy * in[l - Nx].y + idxy * in[l - Nx + 1].y;
                        if (j + 1 < g.Ny) hi = idy2 * in[l + Nx].x + idxy * own[k].y - idxy * in[l + 1].y;
                        nx.x = (cab[k].x + dtrk * dd * (lo + hi)) * inv_da;
                        rmax = fmax(rmax, fabs((double)(nx.x - own[k].x)));
                    }
                    if (j < g.Ny - 1) {
                        R rh, dd, lo = 0, hi = 0;
                        bnd_b(i, rh, dd);
                        if (i > 0) lo = idx2 * in[l - 1].y - idxy * in[l - 1].x + idxy * in[l - 1 + Nx].x;
                        if (i + 1 < Nx) hi = idx2 * in[l + 1].y + idxy * own[k].x - idxy * in[l + Nx].x;
                        nx.y = (cab[k].y + dtrk * dd * (lo + hi)) * inv_db;
                        rmax = fmax(rmax, fabs((double)(nx.y - own[k].y)));
                    }
                    put(out, olo, ohi, k, nx);
                    own[k] = nx;
                }
                const double r = finish_sweep(rmax);
                if (ts_stop(r, A.stop_A, fp32)) { res = s + 1; break; }
            }
            nA += res;
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                if (!ok[k]) continue;
                if (ci[k] < Nx - 1) A.a[off[k]] = own[k].x;
                if (cj[k] < g.Ny - 1) A.b[off[k]] = own[k].y;
            }
            rand_t += 1u;                                    // td.py:313
            cl.sync();
        }
    }
    if (rank == 0 && tid == 0) { A.out[0] = npsi; A.out[1] = nA; A.out[2] = (long long)rand_t; }
}

template <typename R, int NPT>
static int ts_launch(svl_ctx *c, const SmallArgs<R> &A, int nctas, bool *handled) {
    typedef typename V2<R>::type C;
    auto kern = k_td_small<R, NPT>;
    const size_t smem = 2 * (size_t)(A.M + 2 * A.g.Nx) * sizeof(C) + 2 * TS_MAX_CTAS * sizeof(unsigned long long);
    if (smem > 200 * 1024) return 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (nctas > 8) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(nctas); lc.blockDim = dim3(TS_THREADS); lc.stream = c->stream; lc.dynamicSmemBytes = smem;
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
    // the fewest CTAs (a power of two up to 16: one cluster) whose share of nodes fits NPT <= 4 nodes per thread, and a
    // share of at least one grid row (the halo of a CTA must lie in its direct neighbours)
    int npt = 0, nctas = 0, M = 0;
    for (int m = 1; m <= TS_MAX_CTAS && !npt; m *= 2) {
        const long share = (N + m - 1) / m;
        if (share > (long)TS_THREADS * TS_MAX_NPT) continue;
        if (m > 1 && share < g.Nx) break;
        nctas = m; M = (int)share; npt = (int)((share + TS_THREADS - 1) / TS_THREADS);
    }
    if (!npt) return 0;
    // more CTAs shorten a sweep as long as every CTA keeps at least one row and a full warp set busy
    while (nctas < TS_MAX_CTAS && (N + 2 * nctas - 1) / (2 * nctas) >= g.Nx && (N + 2 * nctas - 1) / (2 * nctas) >= TS_THREADS) {
        nctas *= 2; M = (int)((N + nctas - 1) / nctas); npt = (M + TS_THREADS - 1) / TS_THREADS;
    }
    SmallArgs<R> A;
    memset(&A, 0, sizeof(A));
    A.g = g; A.Nt = Nt; A.solveA = solveA; A.M = M;
    A.dt = (R)dt; A.eps = (R)eps; A.kappa2 = (R)kappa2; A.rho = (R)rho; A.H = (R)H; A.lang_psi = (R)lang_psi; A.lang_A = (R)lang_A;
    A.stop_psi = stop_psi; A.stop_A = stop_A;
    A.epsf = epsf ? (const R *)epsf->p[0] : nullptr;
    A.nf = c->nf;
    A.psi = (C *)psi->p[0]; A.a = (R *)ab->p[0]; A.b = (R *)ab->p[1];
    A.rand_t = *rand_t;
    long long *dout = (long long *)(c->d_result + 48);
    A.out = dout;
    switch (npt) {
        case 1: SVL_TRY((ts_launch<R, 1>(c, A, nctas, handled))); break;
        case 2: SVL_TRY((ts_launch<R, 2>(c, A, nctas, handled))); break;
        case 3: SVL_TRY((ts_launch<R, 3>(c, A, nctas, handled))); break;
        default: SVL_TRY((ts_launch<R, 4>(c, A, nctas, handled))); break;
    }
    if (!*handled) return 0;
    c->stat_launches += 1;
    long long *hout = (long long *)(c->h_result + 48);
    SVL_CHECK(cudaMemcpyAsync(hout, dout, 3 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    *rand_t = (uint32_t)hout[2];
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
