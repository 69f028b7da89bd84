// One CG iteration in TWO passes over HBM -- SURVEY.md 8d's fused lower bound, 36R+2 bytes per node
// (as written in svirl/solvers/cg.py:477-544: 84R+4; the three-pass scheme of cg_fused.cu: 40R+3):
//
//   pass A  k_cgp_a   psi <- psi + a_psi d_psi, A <- A + a_A d_A  (utils.h:74-92) + free energy of the new state
//                     (observables.h:251-362) + Jacobians dG/dpsi, dG/dA AT THE NEW STATE (cg.h:16-301), written over
//                     the previous gradient, + the four Polak-Ribiere sums against it (utils.h:13-70)
//                     reads psi,d (4R) a,b,da,db (4R) g_prev (4R) flags (1); writes psi,a,b (4R) g (4R)      = 20R+1
//   pass B  k_cgp_b   d <- beta d - g (beta read on the device, utils.h:97-114) + the 5 / 17 line-search
//                     coefficients of G(psi + a_psi d_psi, A + a_A d_A) (cg.h:315-731)
//                     reads psi (2R) a,b (2R) g (4R) d_old (4R) flags (1); writes d (4R)                     = 16R+1
//
// The update of iteration i and the gradient of iteration i+1 see the same state, so they share one pass; the host
// line search (svirl/solvers/cg.py:227-235, 378-419) sits between B and A exactly as in the reference.
//
// Data movement (the point of this file; the three-pass kernels were latency bound at 0.4-0.6 of the HBM peak with
// per-thread loads two rows ahead): a CTA owns a strip of <= 217 columns x L rows.  A producer warp streams the strip
// row by row into a ring of shared-memory stages with cp.async.bulk (TMA, one bulk copy per plane and row, completion
// on an mbarrier): up to NS-1 rows of every plane are in flight per CTA without costing a register.  Seven consumer warps
// walk down the rows: a lane owns one column, keeps the rows y-1, y, y+1 of the (updated) state in registers (N/S
// neighbours), takes the E/W neighbours from the stage, and evaluates each link variable exp(-i d A) once (the W link
// comes from the neighbouring lane by shuffle; warps overlap by one column so no lane needs a second sincos).
// Reductions accumulate in double per thread over the whole strip and are reduced once per CTA in a fixed order.
#include "common.cuh"

#define CGP_WARPS 7                        // consumer warps; + 1 producer warp = 256 threads, 2 CTAs x 128 registers per SM
#define CGP_CONS (32 * CGP_WARPS)
#define CGP_THREADS (CGP_CONS + 32)        // + one producer warp
#define CGP_MAXPL 9
#define CGP_FULL 0xffffffffu

struct PipeGeom {
    Geo g;
    int nplanes;
    const unsigned char *base[CGP_MAXPL];   // plane base pointers
    int esize[CGP_MAXPL];                   // bytes per element
    int off[CGP_MAXPL];                     // byte offset of the plane's row window inside a stage
    int stage_bytes;                        // multiple of 128
    int HX, NC;                             // halo columns each side; columns per row window (NC = LOUT*8 + 2*HX)
    int WS, nstrips;                        // output columns per strip (<= LOUT*8, multiple of 4), strips per row chunk
    int L;                                  // rows per chunk
    int ylo, yhi;                           // rows produced: [ylo, yhi)
};

__device__ __forceinline__ uint32_t cgp_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cgp_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(cgp_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cgp_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cgp_u32(bar)) : "memory");
}

// Producer warp: rows [r_first, r_last] of the strip starting at column x0, one stage per row.
template <int NS>
__device__ __forceinline__ void cgp_produce(const PipeGeom &G, unsigned char *stages, uint64_t *full, uint64_t *empty,
                                            int x0, int r_first, int r_last) {
    const int lane = threadIdx.x & 31;
    const Geo &g = G.g;
    // column window [x0 - HX, x0 - HX + NC), clamped at the left edge of the plane (the halo columns of the first
    // strip are zero-weight neighbours; they are zeroed once by the consumers)
    const int cut = x0 - G.HX < 0 ? G.HX - x0 : 0;        // columns dropped on the left (0 or HX)
    const int xs = x0 - G.HX + cut;
    uint32_t total = 0;
    for (int p = 0; p < G.nplanes; p++) total += (uint32_t)((G.NC - cut) * G.esize[p]);
    unsigned it = 0;
    for (int r = r_first; r <= r_last; r++, it++) {
        const unsigned s = it & (NS - 1);
        if (it >= NS) cgp_wait(&empty[s], (it / NS - 1) & 1);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cgp_u32(&full[s])), "r"(total) : "memory");
        __syncwarp();
        if (lane < G.nplanes) {
            const int es = G.esize[lane];
            const unsigned char *src = G.base[lane] + ((size_t)(r - g.rb) * g.P + xs) * es;
            unsigned char *dst = stages + (size_t)s * G.stage_bytes + G.off[lane] + cut * es;
            const uint32_t bytes = (uint32_t)((G.NC - cut) * es);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(cgp_u32(dst)), "l"(src), "r"(bytes), "r"(cgp_u32(&full[s])) : "memory");
        }
    }
}

// psi1 * U(ph) - psi0 with (s, c) = sincos(ph)   (cg.h:305-311)
template <typename R, typename C> __device__ __forceinline__ C cgp_gradc(C p0, R s, R c, C p1) {
    C z;
    z.x = p1.x * c + p1.y * s - p0.x;
    z.y = p1.y * c - p1.x * s - p0.y;
    return z;
}

template <typename R> struct CgpState {
    R kappa2, eps, H;
    const R *epsf;
    const uint8_t *nf;
};

// Stage layout, known at compile time (the consumers address every plane as base + immediate): row windows of NC
// columns, each rounded up to 128 bytes, in the order the host adds the planes (cgp_add).
template <typename R, int LOUT> struct PipeDims {
    static constexpr int HX = sizeof(R) == 8 ? 2 : 4;
    static constexpr int NC = (LOUT * CGP_WARPS + 2 * HX + 3) / 4 * 4;
    static constexpr int RB = (NC * (int)sizeof(R) + 127) / 128 * 128;          // real plane
    static constexpr int CB = (NC * 2 * (int)sizeof(R) + 127) / 128 * 128;      // complex plane
};
template <typename R, bool HAVEA, bool SOLVEA, bool PREVPL> struct LayoutA {     // psi, d, [a, b], [da, db], [g, ga, gb]
    typedef PipeDims<R, 31> D;
    static constexpr int o_psi = 0, o_d = D::CB, o_a = 2 * D::CB, o_b = o_a + (HAVEA ? D::RB : 0);
    static constexpr int o_da = o_b + (HAVEA ? D::RB : 0), o_db = o_da + (SOLVEA ? D::RB : 0);
    static constexpr int o_g = o_db + (SOLVEA ? D::RB : 0), o_ga = o_g + (PREVPL ? D::CB : 0);
    static constexpr int o_gb = o_ga + (PREVPL && SOLVEA ? D::RB : 0), stage = o_gb + (PREVPL && SOLVEA ? D::RB : 0);
};
template <typename R, bool HAVEA, bool SOLVEA> struct LayoutB {                  // psi, g, d, [a, b], [ga, gb, da, db]
    typedef PipeDims<R, 32> D;
    static constexpr int o_psi = 0, o_g = D::CB, o_d = 2 * D::CB, o_a = 3 * D::CB, o_b = o_a + (HAVEA ? D::RB : 0);
    static constexpr int o_ga = o_b + (HAVEA ? D::RB : 0), o_gb = o_ga + (SOLVEA ? D::RB : 0);
    static constexpr int o_da = o_gb + (SOLVEA ? D::RB : 0), o_db = o_da + (SOLVEA ? D::RB : 0);
    static constexpr int stage = o_db + (SOLVEA ? D::RB : 0);
};

// Per-flag weight table (16 entries, indexed by the node flag nibble): the DU weights of common.h:13 / cg.h:59-64
// pre-multiplied by the constants they always meet.  w is 0, 1/2 or 1, so every product below is an exact scaling of the
// grid constant and the kernels round exactly like the term-by-term expressions of the reference.
enum { W_G = 0, W_G2, W_CW, W_CE, W_CS, W_CN, W_EE, W_EN, W_JE, W_JN, W_E3, W_N3, W_E12, W_N12, W_N };
template <typename R> __device__ __forceinline__ void cgp_fill_lut(R *lut, const Geo &g) {
    const int f = threadIdx.x;
    if (f >= 16) return;
    const R idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    const R mm = (f & NF_MM) ? (R)1 : (R)0, mp = (f & NF_MP) ? (R)1 : (R)0, pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0;
    const R wW = (R)0.5 * (mm + mp), wE = (R)0.5 * (pm + pp), wS = (R)0.5 * (mm + pm), wN = (R)0.5 * (mp + pp);
    const R gw = (R)0.25 * (wW + wE + wS + wN);
    R *t = lut + f * W_N;
    t[W_G] = gw; t[W_G2] = (R)2.0 * gw;
    t[W_CW] = (R)-2.0 * (wW * idx2); t[W_CE] = (R)-2.0 * (wE * idx2); t[W_CS] = (R)-2.0 * (wS * idy2); t[W_CN] = (R)-2.0 * (wN * idy2);
    t[W_EE] = wE * idx2; t[W_EN] = wN * idy2;
    t[W_JE] = -wE * idx; t[W_JN] = -wN * idy;
    t[W_E3] = -wE * (idx2 / (R)3.0); t[W_N3] = -wN * (idy2 / (R)3.0);
    t[W_E12] = -wE * (idx2 / (R)12.0); t[W_N12] = -wN * (idy2 / (R)12.0);
}

// the two link variables of a node, branch-free; the out-of-range fallback (|phase| > 1e5) is taken by nobody in practice
template <typename R> __device__ __forceinline__ void cgp_sincos2(R pa, R pb, R &sa, R &ca, R &sb, R &cb) {
    sincos_fast(pa, &sa, &ca);
    sincos_fast(pb, &sb, &cb);
    if (!(sincos_fast_ok(pa) && sincos_fast_ok(pb))) { sincos_any(pa, &sa, &ca); sincos_any(pb, &sb, &cb); }
}

__device__ __forceinline__ void cgp_init_barriers(uint64_t *full, uint64_t *empty, int ns) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cgp_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cgp_u32(&empty[s])), "r"(CGP_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// ============================================================================================ pass A
template <typename R, bool HAVEA, bool SOLVEA, bool UPDATE, bool GRAD, bool PREV, int NS>
__global__ void __launch_bounds__(CGP_THREADS, 2)
k_cgp_a(const __grid_constant__ PipeGeom G, CgpState<R> S, R alpha_psi, R alpha_A,
        typename V2<R>::type *__restrict__ psi_out, R *__restrict__ a_out, R *__restrict__ b_out,
        typename V2<R>::type *__restrict__ gpsi, R *__restrict__ ga, R *__restrict__ gb, double *partials) {
    typedef typename V2<R>::type C;
    typedef LayoutA<R, HAVEA, SOLVEA, GRAD && PREV> LY;
    constexpr int HX = LY::D::HX;
    static_assert((NS & (NS - 1)) == 0, "NS must be a power of two");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ R lut[16 * W_N];
    uint64_t *full = (uint64_t *)smem, *empty = full + NS;
    unsigned char *stages = smem + 128;
    const Geo &g = G.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cs = blockIdx.x % G.nstrips, rc = blockIdx.x / G.nstrips;
    const int x0 = cs * G.WS;
    const int ys = G.ylo + rc * G.L, ye = ys + G.L < G.yhi ? ys + G.L : G.yhi;

    cgp_init_barriers(full, empty, NS);
    cgp_fill_lut<R>(lut, g);
    // first strip: the left halo columns are never filled (clamped copies); make them finite once
    if (x0 - HX < 0) {
        for (int q = tid; q < NS * G.nplanes * 8; q += CGP_THREADS) {       // HX * esize <= 64 bytes = 8 doubles
            const int s = q / (G.nplanes * 8), p = (q / 8) % G.nplanes, e = q % 8;
            if (e * 8 < HX * G.esize[p]) ((double *)(stages + (size_t)s * LY::stage + G.off[p]))[e] = 0.0;
        }
    }
    __syncthreads();

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};           // energy, PR sums psi (num, den), A (num, den)
    if (warp == CGP_WARPS) {
        cgp_produce<NS>(G, stages, full, empty, x0, ys - 1, ye);
    } else {
        const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idy2 = (R)g.idy2, idx2 = (R)g.idx2, idxy = (R)g.idxy;
        const int lc = warp * 31 + lane - 1;              // column inside the strip (-1: helper column of the first warp)
        const int col = x0 + lc;
        const bool xin = col >= 0 && col < g.Nx;
        const bool outl = lane >= 1 && lc < G.WS && col < g.Nx;
        // per-thread bases into stage 0: complex planes at tc, real planes at tr; a stage adds pos * LY::stage
        const unsigned char *tc = stages + (size_t)(lc + HX) * sizeof(C), *tr = stages + (size_t)(lc + HX) * sizeof(R);
        auto LC = [&](unsigned pos, int off, int dc) { return *(const C *)(tc + (pos & (NS - 1)) * LY::stage + off + dc * (int)sizeof(C)); };
        auto LR = [&](unsigned pos, int off, int dc) { return *(const R *)(tr + (pos & (NS - 1)) * LY::stage + off + dc * (int)sizeof(R)); };
        // column-dependent pieces of the b-edge stencil (boundary doubling of quirk Q10, cg.h:240-282)
        const bool hasW = col > 0, hasE = col + 1 < g.Nx;
        const R ddb = (col == 0 || col + 1 == g.Nx) ? (R)2 : (R)1;
        const R rhb = col == 0 ? (R)2.0 * S.H * idx : (col + 1 == g.Nx ? (R)-2.0 * S.H * idx : (R)0);
        const R dxdy = dx * dy, dxdy2 = (R)2.0 * dx * dy;
        // running element offset of (col, y) in the pitched planes
        size_t n = g.at(xin ? col : 0, ys - 1);
        const bool mag = S.kappa2 > (R)0;

        struct Row { C p; R a, b, aW; unsigned f; R eps; };
        auto build = [&](unsigned pos, size_t nn) {       // the (updated) state of this lane's node in the row at `pos`
            Row w;
            w.p = LC(pos, LY::o_psi, 0);
            if (UPDATE) {
                const C d = LC(pos, LY::o_d, 0);
                w.p.x = alpha_psi * d.x + w.p.x; w.p.y = alpha_psi * d.y + w.p.y;        // axpy_c (utils.h:74-82)
            }
            w.a = 0; w.b = 0; w.aW = 0;
            if (HAVEA) {
                w.a = LR(pos, LY::o_a, 0); w.b = LR(pos, LY::o_b, 0); w.aW = LR(pos, LY::o_a, -1);
                if (SOLVEA && UPDATE) {
                    w.a = alpha_A * LR(pos, LY::o_da, 0) + w.a;
                    w.b = alpha_A * LR(pos, LY::o_db, 0) + w.b;
                    w.aW = alpha_A * LR(pos, LY::o_da, -1) + w.aW;
                }
            }
            w.f = 0; w.eps = S.eps;
            if (xin) {
                w.f = S.nf[nn];
                if (S.epsf) w.eps = S.epsf[nn];
            } else { w.p.x = 0; w.p.y = 0; w.a = 0; w.b = 0; }
            if (!hasW) w.aW = 0;
            return w;
        };
        // ---- prologue: row ys-1 (S neighbour of the first row) and row ys
        cgp_wait(&full[0], 0);
        Row prv = build(0u, n);
        R sbP, cbP, bEP = 0;
        {
            R t0, t1;
            cgp_sincos2<R>((R)0, dy * prv.b, t0, t1, sbP, cbP);
        }
        if (HAVEA) {
            bEP = LR(0u, LY::o_b, 1);
            if (SOLVEA && UPDATE) bEP = alpha_A * LR(0u, LY::o_db, 1) + bEP;
        }
        __syncwarp();
        if (lane == 0) cgp_arrive(&empty[0]);
        n += g.P;
        cgp_wait(&full[1 & (NS - 1)], (1 / NS) & 1);
        Row cur = build(1u, n);
        unsigned pos = 1;
        for (int y = ys; y < ye; y++, pos++, n += g.P) {
            // `pos` = ring position of row y; row y+1 sits at pos+1
            cgp_wait(&full[(pos + 1) & (NS - 1)], ((pos + 1) / NS) & 1);
            const Row nxt = build(pos + 1, n + g.P);
            // E / W neighbours of row y (updated on the fly from the raw stage values)
            C pE = LC(pos, LY::o_psi, 1), pW = LC(pos, LY::o_psi, -1);
            if (UPDATE) {
                const C dE = LC(pos, LY::o_d, 1), dW = LC(pos, LY::o_d, -1);
                pE.x = alpha_psi * dE.x + pE.x; pE.y = alpha_psi * dE.y + pE.y;
                pW.x = alpha_psi * dW.x + pW.x; pW.y = alpha_psi * dW.y + pW.y;
            }
            R bE = 0, bW = 0;
            if (HAVEA) {
                bE = LR(pos, LY::o_b, 1); bW = LR(pos, LY::o_b, -1);
                if (SOLVEA && UPDATE) {
                    bE = alpha_A * LR(pos, LY::o_db, 1) + bE;
                    bW = alpha_A * LR(pos, LY::o_db, -1) + bW;
                }
            }
            // previous gradient of this node (PR sums)
            C q; q.x = 0; q.y = 0;
            R qa = 0, qb = 0;
            if (GRAD && PREV) {
                q = LC(pos, LY::o_g, 0);
                if (SOLVEA) { qa = LR(pos, LY::o_ga, 0); qb = LR(pos, LY::o_gb, 0); }
            }
            // link variables of this node's E and N links (evaluated whether or not the link carries weight: the weight
            // table zeroes what must not count); the W link is the neighbouring lane's E link
            R sa, ca, sb, cb;
            cgp_sincos2<R>(dx * cur.a, dy * cur.b, sa, ca, sb, cb);
            const R sW = __shfl_up_sync(CGP_FULL, sa, 1), cW = __shfl_up_sync(CGP_FULL, ca, 1);
            if (outl) {
                const R *wt = lut + cur.f * W_N;
                const C p0 = cur.p;
                const R p2 = p0.x * p0.x + p0.y * p0.y;
                // g_grad_jac_psi(psi0, ph, psi1) = 2 (psi0 - psi1 U(ph))  (cg.h:5-12); W and S links enter with the
                // opposite phase: sincos(-x) = (-sin x, cos x)
                const C zW = cgp_gradc<R, C>(p0, -sW, cW, pW), zE = cgp_gradc<R, C>(p0, sa, ca, pE);
                const C zS = cgp_gradc<R, C>(p0, -sbP, cbP, prv.p), zN = cgp_gradc<R, C>(p0, sb, cb, nxt.p);
                if (UPDATE) {
                    R e = wt[W_G] * ((R)0.5 * p2 - cur.eps) * p2;
                    e += wt[W_EE] * (zE.x * zE.x + zE.y * zE.y);
                    e += wt[W_EN] * (zN.x * zN.x + zN.y * zN.y);
                    if (mag && col < g.Nx - 1 && y < g.Ny - 1) {
                        R dB = -S.H;
                        if (HAVEA) dB += idx * (bE - cur.b) - idy * (nxt.a - cur.a);
                        e += S.kappa2 * dB * dB;
                    }
                    acc[0] += (double)e;
                    psi_out[n] = p0;
                    if (SOLVEA) { a_out[n] = cur.a; b_out[n] = cur.b; }
                }
                if (GRAD) {
                    const R pl = wt[W_G2] * (p2 - cur.eps);
                    C gj;
                    gj.x = pl * p0.x; gj.y = pl * p0.y;
                    gj.x += wt[W_CW] * zW.x; gj.y += wt[W_CW] * zW.y;
                    gj.x += wt[W_CE] * zE.x; gj.y += wt[W_CE] * zE.y;
                    gj.x += wt[W_CS] * zS.x; gj.y += wt[W_CS] * zS.y;
                    gj.x += wt[W_CN] * zN.x; gj.y += wt[W_CN] * zN.y;
                    gj.x *= dxdy; gj.y *= dxdy;
                    gpsi[n] = gj;
                    if (PREV) {
                        acc[1] += (double)(gj.x * (gj.x - q.x) + gj.y * (gj.y - q.y));
                        acc[2] += (double)(q.x * q.x + q.y * q.y);
                    }
                    if (SOLVEA) {
                        if (col < g.Nx - 1) {      // dG/da on the a-edge (col, y): curl-curl with quirk Q10 (cg.h:176-217)
                            const bool lo = y == 0, hi = y + 1 == g.Ny;
                            const R dd = (lo || hi) ? (R)2 : (R)1;
                            R v = lo ? (R)-2.0 * S.H * idy : (hi ? (R)2.0 * S.H * idy : (R)0);
                            v += (R)2.0 * idy2 * cur.a;
                            v += lo ? (R)0 : dd * (-idy2 * prv.a + idxy * prv.b - idxy * bEP);
                            v += hi ? (R)0 : dd * (-idy2 * nxt.a - idxy * cur.b + idxy * bE);
                            const R js = (p0.x * pE.y - p0.y * pE.x) * ca - (p0.x * pE.x + p0.y * pE.y) * sa;
                            R w = S.kappa2 * v;
                            w += wt[W_JE] * js;
                            w = dxdy2 * w;
                            ga[n] = w;
                            if (PREV) { acc[3] += (double)(w * (w - qa)); acc[4] += (double)(qa * qa); }
                        }
                        if (y < g.Ny - 1) {        // dG/db on the b-edge (col, y) (cg.h:240-282)
                            R v = rhb;
                            v += (R)2.0 * idx2 * cur.b;
                            v += hasW ? ddb * (-idx2 * bW + idxy * cur.aW - idxy * nxt.aW) : (R)0;
                            v += hasE ? ddb * (-idx2 * bE - idxy * cur.a + idxy * nxt.a) : (R)0;
                            const R js = (p0.x * nxt.p.y - p0.y * nxt.p.x) * cb - (p0.x * nxt.p.x + p0.y * nxt.p.y) * sb;
                            R w = S.kappa2 * v;
                            w += wt[W_JN] * js;
                            w = dxdy2 * w;
                            gb[n] = w;
                            if (PREV) { acc[3] += (double)(w * (w - qb)); acc[4] += (double)(qb * qb); }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) cgp_arrive(&empty[pos & (NS - 1)]);      // row y's stage is free
            prv = cur; sbP = sb; cbP = cb; bEP = bE;
            cur = nxt;
        }
        __syncwarp();
        if (lane == 0) cgp_arrive(&empty[pos & (NS - 1)]);
    }
    block_sum_to_partials<5>(acc, partials, blockIdx.x);
}

// ============================================================================================ pass B
// NV = 5: c0..c4 (cg.h:400-467); NV = 17: c00..c04, c10..c14, c20..c24, c30, c40 (cg.h:528-701).
// Quirk Q11: the coefficient kernels use the scalar eps only.
template <typename R, int NV, bool HAVEA, int NS>
__global__ void __launch_bounds__(CGP_THREADS, 2)
k_cgp_b(const __grid_constant__ PipeGeom G, CgpState<R> S, const double *__restrict__ beta,
        typename V2<R>::type *__restrict__ dpsi_new, R *__restrict__ da_new, R *__restrict__ db_new, double *partials) {
    typedef typename V2<R>::type C;
    constexpr bool SOLVEA = NV == 17;
    typedef LayoutB<R, HAVEA, SOLVEA> LY;
    constexpr int HX = LY::D::HX;
    static_assert((NS & (NS - 1)) == 0, "NS must be a power of two");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ R lut[16 * W_N];
    uint64_t *full = (uint64_t *)smem, *empty = full + NS;
    unsigned char *stages = smem + 128;
    const Geo &g = G.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cs = blockIdx.x % G.nstrips, rc = blockIdx.x / G.nstrips;
    const int x0 = cs * G.WS;
    const int ys = G.ylo + rc * G.L, ye = ys + G.L < G.yhi ? ys + G.L : G.yhi;

    cgp_init_barriers(full, empty, NS);
    cgp_fill_lut<R>(lut, g);
    __syncthreads();

    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = 0.0;
    if (warp == CGP_WARPS) {
        cgp_produce<NS>(G, stages, full, empty, x0, ys, ye);
    } else {
        const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy;
        const R beta_psi = (R)beta[0], beta_A = (R)beta[1];
        constexpr int C1 = NV == 17 ? 5 : 1, C2 = NV == 17 ? 10 : 2, C3 = NV == 17 ? 15 : 3, C4 = NV == 17 ? 16 : 4;
        const int lc = warp * 32 + lane;                  // no W neighbours here: warps do not overlap
        const int col = x0 + lc;
        const bool xin = col < g.Nx;
        const bool outl = lc < G.WS && col < g.Nx;
        const unsigned char *tc = stages + (size_t)(lc + HX) * sizeof(C), *tr = stages + (size_t)(lc + HX) * sizeof(R);
        auto LC = [&](unsigned pos, int off, int dc) { return *(const C *)(tc + (pos & (NS - 1)) * LY::stage + off + dc * (int)sizeof(C)); };
        auto LR = [&](unsigned pos, int off, int dc) { return *(const R *)(tr + (pos & (NS - 1)) * LY::stage + off + dc * (int)sizeof(R)); };
        size_t n = g.at(xin ? col : 0, ys);
        const bool mag = S.kappa2 > (R)0;
        struct Row { C p, d; R a, b, da, db; unsigned f; };
        auto build = [&](unsigned pos, size_t nn) {
            Row w;
            w.p = LC(pos, LY::o_psi, 0);
            const C gg = LC(pos, LY::o_g, 0), dd = LC(pos, LY::o_d, 0);
            w.d.x = beta_psi * dd.x - gg.x; w.d.y = beta_psi * dd.y - gg.y;                 // axmy_c (utils.h:97-104)
            w.a = 0; w.b = 0; w.da = 0; w.db = 0;
            if (HAVEA) { w.a = LR(pos, LY::o_a, 0); w.b = LR(pos, LY::o_b, 0); }
            if (SOLVEA) {
                w.da = beta_A * LR(pos, LY::o_da, 0) - LR(pos, LY::o_ga, 0);
                w.db = beta_A * LR(pos, LY::o_db, 0) - LR(pos, LY::o_gb, 0);
            }
            w.f = 0;
            if (xin) w.f = S.nf[nn];
            else { w.p.x = 0; w.p.y = 0; w.d.x = 0; w.d.y = 0; w.a = 0; w.b = 0; w.da = 0; w.db = 0; }
            return w;
        };
        cgp_wait(&full[0], 0);
        Row cur = build(0u, n);
        unsigned pos = 0;
        for (int y = ys; y < ye; y++, pos++, n += g.P) {
            cgp_wait(&full[(pos + 1) & (NS - 1)], ((pos + 1) / NS) & 1);
            const Row nxt = build(pos + 1, n + g.P);
            const C pE = LC(pos, LY::o_psi, 1);
            C dE;
            {
                const C gg = LC(pos, LY::o_g, 1), dd = LC(pos, LY::o_d, 1);
                dE.x = beta_psi * dd.x - gg.x; dE.y = beta_psi * dd.y - gg.y;
            }
            R bE = 0, dbE = 0;
            if (HAVEA) bE = LR(pos, LY::o_b, 1);
            if (SOLVEA) dbE = beta_A * LR(pos, LY::o_db, 1) - LR(pos, LY::o_gb, 1);
            R sE, cE, sN, cN;
            cgp_sincos2<R>(dx * cur.a, dy * cur.b, sE, cE, sN, cN);
            if (outl) {
                const R *wt = lut + cur.f * W_N;
                const C p0 = cur.p, d0 = cur.d;
                const R gw = wt[W_G];
                const R p2 = p0.x * p0.x + p0.y * p0.y, d2 = d0.x * d0.x + d0.y * d0.y;
                const R tw = (R)2.0 * (p0.x * d0.x + p0.y * d0.y);
                v[0] += (double)(gw * ((R)0.5 * p2 - S.eps) * p2);
                v[C1] += (double)(gw * tw * (p2 - S.eps));
                v[C2] += (double)(gw * (-S.eps * d2 + (R)0.5 * tw * tw + p2 * d2));
                v[C3] += (double)(gw * tw * d2);
                v[C4] += (double)(gw * (R)0.5 * d2 * d2);
#pragma unroll
                for (int dir = 0; dir < 2; dir++) {
                    const R wi2 = wt[dir == 0 ? W_EE : W_EN];                 // w * i2
                    const R s = dir == 0 ? sE : sN, cc = dir == 0 ? cE : cN;
                    const C p1 = dir == 0 ? pE : nxt.p, d1 = dir == 0 ? dE : nxt.d;
                    const C zp = cgp_gradc<R, C>(p0, s, cc, p1), zd = cgp_gradc<R, C>(d0, s, cc, d1);
                    v[0] += (double)(wi2 * (zp.x * zp.x + zp.y * zp.y));
                    v[C1] += (double)(wi2 * (R)2.0 * (zp.x * zd.x + zp.y * zd.y));
                    v[C2] += (double)(wi2 * (zd.x * zd.x + zd.y * zd.y));
                    if (NV == 17) {
                        const R w3 = wt[dir == 0 ? W_E3 : W_N3], w12 = wt[dir == 0 ? W_E12 : W_N12];    // -w*i2/3, -w*i2/12
                        const R dph = (dir == 0 ? dx : dy) * (dir == 0 ? cur.da : cur.db);
                        const R dph2 = dph * dph;
                        // z = x0 * U(-ph) * conj(x1), U(-ph) = c + i s
#define CGP_ZMUL(x0_, x1_, zr, zi)                                           \
    {                                                                        \
        R ur = x0_.x * cc - x0_.y * s, ui = x0_.x * s + x0_.y * cc;          \
        zr = ur * x1_.x + ui * x1_.y;                                        \
        zi = ui * x1_.x - ur * x1_.y;                                        \
    }
                        R zr, zi, z2r, z2i;
                        CGP_ZMUL(p0, p1, zr, zi);
                        v[1] += (double)(wi2 * (R)2.0 * zi * dph);
                        v[2] += (double)(wi2 * zr * dph2);
                        v[3] += (double)(w3 * zi * dph2 * dph);
                        v[4] += (double)(w12 * zr * dph2 * dph2);
                        CGP_ZMUL(p0, d1, zr, zi);
                        CGP_ZMUL(d0, p1, z2r, z2i);
                        zr += z2r; zi += z2i;
                        v[6] += (double)(wi2 * (R)2.0 * zi * dph);
                        v[7] += (double)(wi2 * zr * dph2);
                        v[8] += (double)(w3 * zi * dph2 * dph);
                        v[9] += (double)(w12 * zr * dph2 * dph2);
                        CGP_ZMUL(d0, d1, zr, zi);
                        v[11] += (double)(wi2 * (R)2.0 * zi * dph);
                        v[12] += (double)(wi2 * zr * dph2);
                        v[13] += (double)(w3 * zi * dph2 * dph);
                        v[14] += (double)(w12 * zr * dph2 * dph2);
#undef CGP_ZMUL
                    }
                }
                if (mag && col < g.Nx - 1 && y < g.Ny - 1) {
                    if (NV == 17) {
                        R BH = -S.H;
                        if (HAVEA) BH += idx * (bE - cur.b) - idy * (nxt.a - cur.a);
                        const R dB = idx * (dbE - cur.db) - idy * (nxt.da - cur.da);
                        v[0] += (double)(S.kappa2 * BH * BH);
                        v[1] += (double)(S.kappa2 * (R)2.0 * BH * dB);
                        v[2] += (double)(S.kappa2 * dB * dB);
                    } else {
                        const R dB = idx * (bE - cur.b) - idy * (nxt.a - cur.a) - S.H;
                        v[0] += (double)(S.kappa2 * dB * dB);
                    }
                }
                dpsi_new[n] = cur.d;
                if (SOLVEA) { da_new[n] = cur.da; db_new[n] = cur.db; }
            }
            __syncwarp();
            if (lane == 0) cgp_arrive(&empty[pos & (NS - 1)]);
            cur = nxt;
        }
        __syncwarp();
        if (lane == 0) cgp_arrive(&empty[pos & (NS - 1)]);
    }
    block_sum_to_partials<NV>(v, partials, blockIdx.x);
}

// beta = max(num/den, 0) in real_t, nan -> 0 (divide_scalars_positive, utils.h:140-146); sums[1..4] of pass A
template <typename R>
__global__ void k_cgp_beta(const double *__restrict__ sums, double *beta) {
    if (threadIdx.x < 2) {
        R q = (R)sums[1 + 2 * threadIdx.x] / (R)sums[2 + 2 * threadIdx.x];
        beta[threadIdx.x] = (q > (R)0) ? (double)q : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <typename R> static constexpr int cgp_ns() { return 4; }

// strips of <= lout*8 output columns (multiple of 4: 16-byte aligned row windows in every plane), row chunks of L rows
static void cgp_geom(svl_ctx *c, PipeGeom &G, int rsize, int lout, int L) {
    memset(&G, 0, sizeof(G));
    G.g = c->g;
    G.HX = rsize == 8 ? 2 : 4;
    const int wmax = lout * CGP_WARPS;
    G.NC = (wmax + 2 * G.HX + 3) / 4 * 4;
    G.nstrips = (c->g.Nx + wmax - 1) / wmax;
    int ws = (c->g.Nx + G.nstrips - 1) / G.nstrips;
    ws = (ws + 3) / 4 * 4;
    if (ws > wmax / 4 * 4) { ws = wmax / 4 * 4; G.nstrips = (c->g.Nx + ws - 1) / ws; }
    G.WS = ws;
    G.ylo = c->g.j0; G.yhi = c->g.j1;
    G.L = L;
}
static void cgp_add(PipeGeom &G, const void *base, int esize) {
    const int k = G.nplanes++;
    G.base[k] = (const unsigned char *)base;
    G.esize[k] = esize;
    G.off[k] = G.stage_bytes;
    G.stage_bytes += (G.NC * esize + 127) / 128 * 128;
}
static int cgp_chunk_rows(const svl_ctx *c, int nstrips) {
    // enough CTAs for a few waves of 2 x 148 resident ones, chunks not shorter than 32 rows (each chunk re-reads 1-2 rows)
    const int rows = c->g.j1 - c->g.j0;
    int L = 64;
    while (L > 16 && (long)nstrips * ((rows + L - 1) / L) < 148 * 2 * 3) L >>= 1;
    return L;
}

template <typename R>
static int cgp_pass_a_t(svl_ctx *c, int solveA, int do_update, int do_grad, int have_prev, double kappa2, double eps,
                        const svl_buf *epsf, double H, svl_buf *psi, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A,
                        double alpha_psi, double alpha_A, svl_buf *g_psi, svl_buf *g_A, double *beta, double *E_out) {
    typedef typename V2<R>::type C;
    constexpr int NS = cgp_ns<R>();
    PipeGeom G;
    cgp_geom(c, G, sizeof(R), 31, 64);
    G.L = cgp_chunk_rows(c, G.nstrips);
    const bool havea = ab != nullptr;
    cgp_add(G, psi->p[0], sizeof(C));
    cgp_add(G, d_psi->p[0], sizeof(C));
    if (havea) { cgp_add(G, ab->p[0], sizeof(R)); cgp_add(G, ab->p[1], sizeof(R)); }
    if (solveA) { cgp_add(G, d_A->p[0], sizeof(R)); cgp_add(G, d_A->p[1], sizeof(R)); }
    if (do_grad && have_prev) {
        cgp_add(G, g_psi->p[0], sizeof(C));
        if (solveA) { cgp_add(G, g_A->p[0], sizeof(R)); cgp_add(G, g_A->p[1], sizeof(R)); }
    }
    const size_t smem = 128 + (size_t)NS * G.stage_bytes;
    const int nb = G.nstrips * ((G.yhi - G.ylo + G.L - 1) / G.L);
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * 5));
    CgpState<R> S;
    S.kappa2 = (R)kappa2; S.eps = (R)eps; S.H = (R)H;
    S.epsf = epsf ? (const R *)epsf->p[0] : nullptr;
    S.nf = c->nf;
    svl_buf *pn = nullptr, *An = nullptr;
    if (do_update) {
        SVL_TRY(svl_scratch_node(c, 0, &pn));
        if (solveA) SVL_TRY(svl_scratch_edge(c, 0, &An));
    }
    C *po = pn ? (C *)pn->p[0] : nullptr;
    R *ao = An ? (R *)An->p[0] : nullptr, *bo = An ? (R *)An->p[1] : nullptr;
    C *gp = (C *)g_psi->p[0];
    R *gA0 = solveA ? (R *)g_A->p[0] : nullptr, *gA1 = solveA ? (R *)g_A->p[1] : nullptr;
#define CGP_A_LAUNCH(HA, SA, UP, GR, PV)                                                                            \
    do {                                                                                                            \
        auto kern = k_cgp_a<R, HA, SA, UP, GR, PV, NS>;                                                             \
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        kern<<<nb, CGP_THREADS, smem, c->stream>>>(G, S, (R)alpha_psi, (R)alpha_A, po, ao, bo, gp, gA0, gA1, c->partials); \
    } while (0)
    const int key = (havea ? 16 : 0) | (solveA ? 8 : 0) | (do_update ? 4 : 0) | (do_grad ? 2 : 0) | ((do_grad && have_prev) ? 1 : 0);
    switch (key) {
        // first call of a cg(): gradient only
        case 16 | 8 | 2: CGP_A_LAUNCH(true, true, false, true, false); break;
        case 16 | 2: CGP_A_LAUNCH(true, false, false, true, false); break;
        case 2: CGP_A_LAUNCH(false, false, false, true, false); break;
        // regular iteration: update + energy + gradient + PR sums
        case 16 | 8 | 4 | 2 | 1: CGP_A_LAUNCH(true, true, true, true, true); break;
        case 16 | 4 | 2 | 1: CGP_A_LAUNCH(true, false, true, true, true); break;
        case 4 | 2 | 1: CGP_A_LAUNCH(false, false, true, true, true); break;
        // last iteration of a cg(): update + energy only
        case 16 | 8 | 4: CGP_A_LAUNCH(true, true, true, false, false); break;
        case 16 | 4: CGP_A_LAUNCH(true, false, true, false, false); break;
        case 4: CGP_A_LAUNCH(false, false, true, false, false); break;
        default: svl_set_error("cg pass A: unsupported mode %d", key); return 2;
    }
#undef CGP_A_LAUNCH
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    if (do_update) {
        SVL_TRY(svl_swap(c, psi, pn));
        if (solveA) SVL_TRY(svl_swap(c, ab, An));
    }
    // second stage of the sums; beta stays on the device for pass B and is mirrored to the host (quirk Q6)
    double *dbeta = c->d_result + 32;
    if (do_update || (do_grad && have_prev)) SVL_TRY(svl_finish_sum(c, nb, 5, 1.0, nullptr));
    if (do_grad && have_prev) {
        k_cgp_beta<R><<<1, 32, 0, c->stream>>>(c->d_result, dbeta);
        SVL_CHECK(cudaGetLastError());
        c->stat_launches += 1;
    } else {
        c->h_result[40] = beta[0]; c->h_result[41] = beta[1];
        SVL_CHECK(cudaMemcpyAsync(dbeta, c->h_result + 40, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (do_update || (do_grad && have_prev)) {
        SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaMemcpyAsync(c->h_result + 32, dbeta, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        if (E_out && do_update) *E_out = c->h_result[0] * (double)((R)c->g.dx * (R)c->g.dy);
        if (do_grad && have_prev) { beta[0] = c->h_result[32]; if (solveA) beta[1] = c->h_result[33]; }
    }
    return 0;
}

template <typename R>
static int cgp_pass_b_t(svl_ctx *c, int solveA, double kappa2, double eps, double H, const svl_buf *psi, const svl_buf *ab,
                        const svl_buf *g_psi, const svl_buf *g_A, svl_buf *d_psi, svl_buf *d_A, double *c_out) {
    typedef typename V2<R>::type C;
    constexpr int NS = cgp_ns<R>();
    PipeGeom G;
    cgp_geom(c, G, sizeof(R), 32, 64);
    G.L = cgp_chunk_rows(c, G.nstrips);
    const bool havea = ab != nullptr;
    cgp_add(G, psi->p[0], sizeof(C));
    cgp_add(G, g_psi->p[0], sizeof(C));
    cgp_add(G, d_psi->p[0], sizeof(C));
    if (havea) { cgp_add(G, ab->p[0], sizeof(R)); cgp_add(G, ab->p[1], sizeof(R)); }
    if (solveA) {
        cgp_add(G, g_A->p[0], sizeof(R)); cgp_add(G, g_A->p[1], sizeof(R));
        cgp_add(G, d_A->p[0], sizeof(R)); cgp_add(G, d_A->p[1], sizeof(R));
    }
    const size_t smem = 128 + (size_t)NS * G.stage_bytes;
    const int nb = G.nstrips * ((G.yhi - G.ylo + G.L - 1) / G.L);
    const int nv = solveA ? 17 : 5;
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * nv));
    CgpState<R> S;
    S.kappa2 = (R)kappa2; S.eps = (R)eps; S.H = (R)H;
    S.epsf = nullptr;                                   // quirk Q11: scalar eps only
    S.nf = c->nf;
    svl_buf *dn_psi = nullptr, *dn_A = nullptr;
    SVL_TRY(svl_scratch_node(c, 0, &dn_psi));
    if (solveA) SVL_TRY(svl_scratch_edge(c, 0, &dn_A));
    double *dbeta = c->d_result + 32;
    C *dpo = (C *)dn_psi->p[0];
    R *dao = dn_A ? (R *)dn_A->p[0] : nullptr, *dbo = dn_A ? (R *)dn_A->p[1] : nullptr;
#define CGP_B_LAUNCH(NVV, HA)                                                                                       \
    do {                                                                                                            \
        auto kern = k_cgp_b<R, NVV, HA, NS>;                                                                        \
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        kern<<<nb, CGP_THREADS, smem, c->stream>>>(G, S, dbeta, dpo, dao, dbo, c->partials);                        \
    } while (0)
    if (solveA) CGP_B_LAUNCH(17, true);
    else if (havea) CGP_B_LAUNCH(5, true);
    else CGP_B_LAUNCH(5, false);
#undef CGP_B_LAUNCH
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    SVL_TRY(svl_swap(c, d_psi, dn_psi));
    if (solveA) SVL_TRY(svl_swap(c, d_A, dn_A));
    return svl_finish_sum(c, nb, nv, (double)((R)c->g.dx * (R)c->g.dy), c_out);      // one host sync
}

static int cgp_check(svl_ctx *c, const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, const svl_buf *d_psi,
                     const svl_buf *g_psi, int solveA, const svl_buf *d_A, const svl_buf *g_A) {
    SVL_REQUIRE(c && psi && d_psi && g_psi, "null argument");
    SVL_REQUIRE(psi->kind == SVL_NODE_C && d_psi->kind == SVL_NODE_C && g_psi->kind == SVL_NODE_C, "psi-side buffers must be SVL_NODE_C");
    SVL_REQUIRE(!ab || ab->kind == SVL_EDGE, "ab must be SVL_EDGE");
    SVL_REQUIRE(!solveA || (ab && d_A && g_A && d_A->kind == SVL_EDGE && g_A->kind == SVL_EDGE), "A-side buffers must be SVL_EDGE");
    SVL_REQUIRE(!abei, "the two-pass CG iteration does not take an external potential (use svl_cg_begin / svl_cg_end)");
    SVL_REQUIRE(!c->slab_on, "the two-pass CG iteration runs on one GPU (row slabs: svl_cg_begin / svl_cg_end)");
    return 0;
}

extern "C" int svl_cg_pass_a(svl_ctx *c, int solveA, int do_update, int do_grad, int have_prev, double kappa2, double eps,
                             const svl_buf *epsf, double H, svl_buf *psi, const svl_buf *abei, svl_buf *ab,
                             const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi, double alpha_A, svl_buf *g_psi,
                             svl_buf *g_A, double *beta, double *E_out) {
    SVL_TRY(cgp_check(c, psi, abei, ab, d_psi, g_psi, solveA, d_A, g_A));
    SVL_REQUIRE(beta && (do_update || do_grad), "nothing to do");
    SVL_REQUIRE(!epsf || epsf->kind == SVL_NODE_R, "eps_field must be SVL_NODE_R");
    if (c->rsize == 4) return cgp_pass_a_t<float>(c, solveA, do_update, do_grad, have_prev, kappa2, eps, epsf, H, psi, ab, d_psi,
                                                  d_A, alpha_psi, alpha_A, g_psi, g_A, beta, E_out);
    return cgp_pass_a_t<double>(c, solveA, do_update, do_grad, have_prev, kappa2, eps, epsf, H, psi, ab, d_psi, d_A, alpha_psi,
                                alpha_A, g_psi, g_A, beta, E_out);
}

extern "C" int svl_cg_pass_b(svl_ctx *c, int solveA, double kappa2, double eps, double H, const svl_buf *psi,
                             const svl_buf *abei, const svl_buf *ab, const svl_buf *g_psi, const svl_buf *g_A, svl_buf *d_psi,
                             svl_buf *d_A, double *c_out) {
    SVL_TRY(cgp_check(c, psi, abei, ab, d_psi, g_psi, solveA, d_A, g_A));
    SVL_REQUIRE(c_out, "null c_out");
    if (c->rsize == 4) return cgp_pass_b_t<float>(c, solveA, kappa2, eps, H, psi, ab, g_psi, g_A, d_psi, d_A, c_out);
    return cgp_pass_b_t<double>(c, solveA, kappa2, eps, H, psi, ab, g_psi, g_A, d_psi, d_A, c_out);
}
