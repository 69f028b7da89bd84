// One CG iteration in three passes over HBM (fused lower bound ~ 40R+3 bytes/node
// against 84R+4 as written in svirl/solvers/cg.py:238-551):
//
//   k_cgf_grad    Jacobians dG/dpsi, dG/dA at (psi, A)  + the four Polak-Ribiere sums
//                 (cg.h:16-301, utils.h:13-70)
//   k_cgf_coef    direction update d <- beta d - g (beta read on the device) fused with the 5 / 17
//                 line-search coefficients of G(psi + a_psi d_psi, A + a_A d_A)
//                 (utils.h:97-114, cg.h:315-731)
//   k_cgf_update  psi <- psi + a_psi d_psi, A <- A + a_A d_A fused with the free energy of the new
//                 state (utils.h:74-92, observables.h:251-362)
//
// All three walk the grid the same way: a warp owns 32 consecutive columns and a strip of V rows;
// the lanes hold one node each, the E/W neighbours come from warp shuffles (adjacent warps overlap by
// one or two columns so that no lane loads a neighbour by itself), the N neighbour is the next row
// of the strip, the S neighbour the previous one; rows are loaded two ahead of their use.  Every link variable exp(-i d A) is evaluated once by the node that owns the edge
// (2 sincos per node instead of 6 in the Jacobians) and never stored.  CTAs are persistent
// (tile = 32 columns x 8V rows, tiles dealt round-robin), reductions accumulate in double in
// registers across all tiles of a CTA and are reduced once per CTA in a fixed order, so the sums are
// run-to-run reproducible and the second stage adds ~1200 partials instead of one per 256 nodes.
// Updated fields are written out of place (a neighbour's old value must stay readable) and the
// caller swaps storage.
#include "common.cuh"

#define CGF_WARPS 8
#define CGF_THREADS (32 * CGF_WARPS)
#define FULL 0xffffffffu

template <typename R> struct CgfState {     // what all three kernels read
    Geo g;
    int V;                                  // rows per strip (run time: 32 on large grids, fewer on small ones)
    int ext_lo, ext_hi;                     // slabs: rows computed beyond [j0, j1) towards a neighbour (sums stay on owned rows)
    R kappa2, eps, H;
    const R *epsf;
    const uint8_t *nf;
    const typename V2<R>::type *psi;
    const R *ae, *be, *a, *b;
};

template <typename C> __device__ __forceinline__ C shfl_down_c(C v) {
    C r;
    r.x = __shfl_down_sync(FULL, v.x, 1);
    r.y = __shfl_down_sync(FULL, v.y, 1);
    return r;
}
template <typename C> __device__ __forceinline__ C shfl_up_c(C v) {
    C r;
    r.x = __shfl_up_sync(FULL, v.x, 1);
    r.y = __shfl_up_sync(FULL, v.y, 1);
    return r;
}
template <typename R> __device__ __forceinline__ void du_w(unsigned f, R &wW, R &wE, R &wS, R &wN, R &gw) {
    R mm = (f & NF_MM) ? (R)1 : (R)0, mp = (f & NF_MP) ? (R)1 : (R)0;
    R pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0;
    wW = (R)0.5 * (mm + mp); wE = (R)0.5 * (pm + pp);
    wS = (R)0.5 * (mm + pm); wN = (R)0.5 * (mp + pp);
    gw = (R)0.25 * (wW + wE + wS + wN);
}
// psi1 * U(ph) - psi0 with (s, c) = sincos(ph)   (cg.h:305-311)
template <typename R, typename C> __device__ __forceinline__ C gradc(C p0, R s, R c, C p1) {
    C z;
    z.x = p1.x * c + p1.y * s - p0.x;
    z.y = p1.y * c - p1.x * s - p0.y;
    return z;
}

// One lane asks the L2 for a whole row segment (cp.async.bulk.prefetch: no registers, no scoreboard):
// at the start of a strip the lanes 0..V-1 request rows ys+2..ye of every plane the kernel reads, so
// the register loads two rows ahead hit L2 (~250 cycles) instead of HBM (~1000 under load).
__device__ __forceinline__ void cgf_pf(const void *base, size_t elem, int nelem, int esize) {
    if (!base) return;
    const size_t b0 = elem * (size_t)esize, a0 = b0 & ~(size_t)15;
    const uint32_t nb = (uint32_t)(((b0 + (size_t)nelem * esize + 15) & ~(size_t)15) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)base + a0), "r"(nb) : "memory");
}

// Tile walk shared by the three kernels: the 8 warps of a CTA sit side by side along x (a CTA row is
// one contiguous 2-4 KB run per plane: DRAM pages stay open) and walk the same V rows.  A warp reads
// 32 consecutive columns but produces only WOUT
// of them (lanes LPAD .. LPAD+WOUT-1): the remaining lane(s) exist to hand their values to the
// neighbours by shuffle, so no lane ever has to load a neighbour column on its own (a divergent,
// fully exposed memory latency per row otherwise).  Rows are loaded two ahead of the one being
// computed, so the loads of row y+2 are in flight while row y is evaluated.
#define CGF_TILE_LOOP_BEGIN(WOUT, LPAD)                                                            \
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;           \
    const int ntx = (g.Nx + (WOUT) * nwarp - 1) / ((WOUT) * nwarp);                                \
    const int ylo_ = g.j0 - S.ext_lo, yhi_ = g.j1 + S.ext_hi;                                      \
    const int nty = (yhi_ - ylo_ + S.V - 1) / S.V;                                                 \
    for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {                                      \
        const int i = ((t % ntx) * nwarp + warp) * (WOUT) + lane - (LPAD);                         \
        const int ys = ylo_ + (t / ntx) * S.V;                                                     \
        const int ye = ys + S.V < yhi_ ? ys + S.V : yhi_;                                          \
        if (i - lane + (LPAD) >= g.Nx) continue;                                                   \
        const bool in = lane >= (LPAD) && lane < (LPAD) + (WOUT) && i < g.Nx;
#define CGF_TILE_LOOP_END }
// inside the tile loop: lane l requests row ys+2+l (<= ye; V <= 32) of the listed planes, columns of this warp
#define CGF_PF_BEGIN                                                                               \
    {                                                                                              \
        const int pfy = ys + 2 + lane;                                                             \
        const int pi0 = (i - lane) < 0 ? 0 : (i - lane);                                           \
        const int pn_ = (pi0 + 33 <= g.Nx ? pi0 + 33 : g.Nx) - pi0;                                \
        if (pfy <= ye && pn_ > 0) {                                                \
            const size_t pe = g.at(pi0, pfy);
#define CGF_PF(ptr, esize) cgf_pf((const void *)(ptr), pe, pn_, (int)(esize));
#define CGF_PF_END }}

// ============================================================================= update + energy
template <typename R> struct RawU { typename V2<R>::type p, d; R a, b, da, db, ea, eb, eps; unsigned f; };
template <typename R> struct RowU { typename V2<R>::type p; R a, b, ea, eb, eps; unsigned f; };

template <typename R, bool SOLVEA, bool EXT>
__global__ void __launch_bounds__(CGF_THREADS, 2)
k_cgf_update(CgfState<R> S, const typename V2<R>::type *__restrict__ dpsi, const R *__restrict__ da,
             const R *__restrict__ db, R alpha_psi, R alpha_A, typename V2<R>::type *__restrict__ psi_out,
             R *__restrict__ a_out, R *__restrict__ b_out, double *partials) {
    typedef typename V2<R>::type C;
    const Geo &g = S.g;
    const R *const ae_ = EXT ? S.ae : nullptr, *const be_ = EXT ? S.be : nullptr;   // compile-time absent without an external potential
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    double acc[1] = {0.0};
    // Raw row data (plain loads, no arithmetic: the loads stay in flight until the row is combined two
    // iterations later) and the updated state of a node; zero outside the grid (rows are always
    // inside the plane).
    auto loadraw = [&](int ii, int y) {
        RawU<R> r;
        r.p.x = 0; r.p.y = 0; r.d.x = 0; r.d.y = 0; r.a = 0; r.b = 0; r.da = 0; r.db = 0; r.ea = 0; r.eb = 0;
        r.eps = S.eps; r.f = 0;
        if (ii >= 0 && ii < g.Nx) {
            const size_t n = g.at(ii, y);
            r.p = S.psi[n]; r.d = dpsi[n];
            r.f = S.nf[n];
            if (S.epsf) r.eps = S.epsf[n];
            if (S.a) {
                r.a = S.a[n]; r.b = S.b[n];
                if (SOLVEA) { r.da = da[n]; r.db = db[n]; }
            }
            if (ae_) { r.ea = ae_[n]; r.eb = be_[n]; }
        }
        return r;
    };
    auto combine = [&](const RawU<R> &w) {
        RowU<R> r;
        r.p.x = alpha_psi * w.d.x + w.p.x; r.p.y = alpha_psi * w.d.y + w.p.y;      // axpy_c (utils.h:74-82)
        r.a = w.a; r.b = w.b;
        if (SOLVEA) { r.a = alpha_A * w.da + w.a; r.b = alpha_A * w.db + w.b; }
        r.ea = w.ea; r.eb = w.eb; r.eps = w.eps; r.f = w.f;
        return r;
    };
    CGF_TILE_LOOP_BEGIN(31, 0)
        CGF_PF_BEGIN
            CGF_PF(S.psi, sizeof(C)) CGF_PF(dpsi, sizeof(C)) CGF_PF(S.a, sizeof(R)) CGF_PF(S.b, sizeof(R))
            CGF_PF(da, sizeof(R)) CGF_PF(db, sizeof(R)) CGF_PF(ae_, sizeof(R)) CGF_PF(be_, sizeof(R))
            CGF_PF(S.epsf, sizeof(R)) CGF_PF(S.nf, 1)
        CGF_PF_END
        RowU<R> cur = combine(loadraw(i, ys));
        RawU<R> rawn = loadraw(i, ys + 1), raw2 = loadraw(ys + 2 <= ye ? i : -1, ys + 2);
        for (int y = ys; y < ye; y++) {
            const RowU<R> nxt = combine(rawn);
            rawn = raw2;
            raw2 = loadraw(y + 3 <= ye ? i : -1, y + 3);
            const C pE = shfl_down_c<C>(cur.p);
            const R bE = __shfl_down_sync(FULL, cur.b, 1), ebE = __shfl_down_sync(FULL, cur.eb, 1);
            if (in) {
                const size_t n = g.at(i, y);
                const unsigned f = cur.f;
                const R eps = cur.eps;
                R e = 0;
                if (f) {
                    R wW, wE, wS, wN, gw;
                    du_w<R>(f, wW, wE, wS, wN, gw);
                    const R p2 = cur.p.x * cur.p.x + cur.p.y * cur.p.y;
                    e += gw * ((R)0.5 * p2 - eps) * p2;
                    R s, c;
                    if (f & (NF_PM | NF_PP)) {
                        sincos_r<R>(dx * (ae_ ? cur.ea + cur.a : cur.a), &s, &c);
                        const C z = gradc<R, C>(cur.p, s, c, pE);
                        e += wE * idx2 * (z.x * z.x + z.y * z.y);
                    }
                    if (f & (NF_MP | NF_PP)) {
                        sincos_r<R>(dy * (ae_ ? cur.eb + cur.b : cur.b), &s, &c);
                        const C z = gradc<R, C>(cur.p, s, c, nxt.p);
                        e += wN * idy2 * (z.x * z.x + z.y * z.y);
                    }
                }
                if (S.kappa2 > (R)0 && i < g.Nx - 1 && y < g.Ny - 1) {
                    R dB = -S.H;
                    if (ae_) dB += idx * (ebE - cur.eb) - idy * (nxt.ea - cur.ea);
                    if (S.a) dB += idx * (bE - cur.b) - idy * (nxt.a - cur.a);
                    e += S.kappa2 * dB * dB;
                }
                if (y >= g.j0 && y < g.j1) acc[0] += (double)e;          // halo rows (slabs) are updated, not summed
                psi_out[n] = cur.p;
                if (SOLVEA) { a_out[n] = cur.a; b_out[n] = cur.b; }
            }
            cur = nxt;
        }
    CGF_TILE_LOOP_END
    block_sum_to_partials<1>(acc, partials, blockIdx.x);
}

// ============================================================================= direction + coefficients
template <typename R> struct RawD { typename V2<R>::type p, d, g; R a, b, ea, eb, da, db, ga, gb; unsigned f; };
template <typename R> struct RowD { typename V2<R>::type p, d; R a, b, ea, eb, da, db; unsigned f; };

// NV = 5: c0..c4 (cg.h:400-467); NV = 17: c00..c04, c10..c14, c20..c24, c30, c40 (cg.h:528-701).
// Quirk Q11: the coefficient kernels use the scalar eps only.
// (the 17-coefficient fp64 variant runs 128-thread CTAs, three per SM: 170 registers per thread hold
// the accumulators and three rows without spilling)
template <typename R, int NV, bool EXT>
__global__ void __launch_bounds__((NV == 17 && sizeof(R) == 8) ? 128 : CGF_THREADS, (NV == 17 && sizeof(R) == 8) ? 3 : 2)
k_cgf_coef(CgfState<R> S, const double *__restrict__ beta, const typename V2<R>::type *__restrict__ gpsi,
           const R *__restrict__ ga, const R *__restrict__ gb, const typename V2<R>::type *__restrict__ dpsi_old,
           const R *__restrict__ da_old, const R *__restrict__ db_old, typename V2<R>::type *__restrict__ dpsi_new,
           R *__restrict__ da_new, R *__restrict__ db_new, double *partials) {
    typedef typename V2<R>::type C;
    constexpr bool SOLVEA = NV == 17;
    const Geo &g = S.g;
    const R *const ae_ = EXT ? S.ae : nullptr, *const be_ = EXT ? S.be : nullptr;   // compile-time absent without an external potential
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    // w is 0, 1/2 or 1, so (-w*i2)/3 == -w*(i2/3) bit for bit: the divisions of cg.h:600-640 leave the node loop
    const R idx2_3 = idx2 / (R)3.0, idy2_3 = idy2 / (R)3.0, idx2_12 = idx2 / (R)12.0, idy2_12 = idy2 / (R)12.0;
    const R beta_psi = (R)beta[0], beta_A = (R)beta[1];
    const int C1 = NV == 17 ? 5 : 1, C2 = NV == 17 ? 10 : 2, C3 = NV == 17 ? 15 : 3, C4 = NV == 17 ? 16 : 4;
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = 0.0;
    auto loadraw = [&](int ii, int y) {
        RawD<R> r;
        r.p.x = 0; r.p.y = 0; r.d.x = 0; r.d.y = 0; r.g.x = 0; r.g.y = 0; r.a = 0; r.b = 0; r.ea = 0; r.eb = 0;
        r.da = 0; r.db = 0; r.ga = 0; r.gb = 0; r.f = 0;
        if (ii >= 0 && ii < g.Nx) {
            const size_t n = g.at(ii, y);
            r.p = S.psi[n]; r.d = dpsi_old[n]; r.g = gpsi[n];
            r.f = S.nf[n];
            if (S.a) { r.a = S.a[n]; r.b = S.b[n]; }
            if (ae_) { r.ea = ae_[n]; r.eb = be_[n]; }
            if (SOLVEA) { r.da = da_old[n]; r.db = db_old[n]; r.ga = ga[n]; r.gb = gb[n]; }
        }
        return r;
    };
    auto combine = [&](const RawD<R> &w) {
        RowD<R> r;
        r.p = w.p;
        r.d.x = beta_psi * w.d.x - w.g.x; r.d.y = beta_psi * w.d.y - w.g.y;        // axmy_c (utils.h:97-104)
        r.a = w.a; r.b = w.b; r.ea = w.ea; r.eb = w.eb; r.f = w.f;
        r.da = 0; r.db = 0;
        if (SOLVEA) { r.da = beta_A * w.da - w.ga; r.db = beta_A * w.db - w.gb; }
        return r;
    };
    CGF_TILE_LOOP_BEGIN(31, 0)
        CGF_PF_BEGIN
            CGF_PF(S.psi, sizeof(C)) CGF_PF(dpsi_old, sizeof(C)) CGF_PF(gpsi, sizeof(C)) CGF_PF(S.a, sizeof(R))
            CGF_PF(S.b, sizeof(R)) CGF_PF(da_old, sizeof(R)) CGF_PF(db_old, sizeof(R)) CGF_PF(ga, sizeof(R))
            CGF_PF(gb, sizeof(R)) CGF_PF(ae_, sizeof(R)) CGF_PF(be_, sizeof(R)) CGF_PF(S.nf, 1)
        CGF_PF_END
        // combine row y+1 first, then refill the raw registers with row y+2: its loads are in flight
        // while row y is evaluated (the 17 accumulators leave room for one raw row only)
        RowD<R> cur = combine(loadraw(i, ys));
        RawD<R> rawn = loadraw(i, ys + 1);
        for (int y = ys; y < ye; y++) {
            const RowD<R> nxt = combine(rawn);
            rawn = loadraw(y + 2 <= ye ? i : -1, y + 2);
            const C pE = shfl_down_c<C>(cur.p), dE = shfl_down_c<C>(cur.d);
            const R bE = __shfl_down_sync(FULL, cur.b, 1), ebE = __shfl_down_sync(FULL, cur.eb, 1);
            const R dbE = __shfl_down_sync(FULL, cur.db, 1);
            if (in) {
                const size_t n = g.at(i, y);
                const unsigned f = (y >= g.j0 && y < g.j1) ? cur.f : 0u;     // halo rows (slabs): direction only, no sums
                const bool own = y >= g.j0 && y < g.j1;
                if (f) {
                    R wW, wE, wS, wN, gw;
                    du_w<R>(f, wW, wE, wS, wN, gw);
                    const C p0 = cur.p, d0 = cur.d;
                    const R p2 = p0.x * p0.x + p0.y * p0.y, d2 = d0.x * d0.x + d0.y * d0.y;
                    const R tw = (R)2.0 * (p0.x * d0.x + p0.y * d0.y);
                    v[0] += (double)(gw * ((R)0.5 * p2 - S.eps) * p2);
                    v[C1] += (double)(gw * tw * (p2 - S.eps));
                    v[C2] += (double)(gw * (-S.eps * d2 + (R)0.5 * tw * tw + p2 * d2));
                    v[C3] += (double)(gw * tw * d2);
                    v[C4] += (double)(gw * (R)0.5 * d2 * d2);
#pragma unroll
                    for (int dir = 0; dir < 2; dir++) {
                        const bool on = dir == 0 ? (f & (NF_PM | NF_PP)) : (f & (NF_MP | NF_PP));
                        if (!on) continue;
                        const R w = dir == 0 ? wE : wN, i2 = dir == 0 ? idx2 : idy2, d = dir == 0 ? dx : dy;
                        // (1/3, 1/12 of the Taylor terms folded into per-direction constants: no division per node)
                        const R i2_3 = dir == 0 ? idx2_3 : idy2_3, i2_12 = dir == 0 ? idx2_12 : idy2_12;
                        R ph = 0;
                        if (ae_) ph += d * (dir == 0 ? cur.ea : cur.eb);
                        if (S.a) ph += d * (dir == 0 ? cur.a : cur.b);
                        R s, c;
                        sincos_r<R>(ph, &s, &c);
                        const C p1 = dir == 0 ? pE : nxt.p, d1 = dir == 0 ? dE : nxt.d;
                        const C zp = gradc<R, C>(p0, s, c, p1), zd = gradc<R, C>(d0, s, c, d1);
                        v[0] += (double)(w * i2 * (zp.x * zp.x + zp.y * zp.y));
                        v[C1] += (double)(w * i2 * (R)2.0 * (zp.x * zd.x + zp.y * zd.y));
                        v[C2] += (double)(w * i2 * (zd.x * zd.x + zd.y * zd.y));
                        if (NV == 17) {
                            const R dph = d * (dir == 0 ? cur.da : cur.db);
                            const R dph2 = dph * dph;
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
                if (own && S.kappa2 > (R)0 && i < g.Nx - 1 && y < g.Ny - 1) {
                    if (NV == 17) {
                        R BH = -S.H;
                        if (ae_) BH += idx * (ebE - cur.eb) - idy * (nxt.ea - cur.ea);
                        if (S.a) BH += idx * (bE - cur.b) - idy * (nxt.a - cur.a);
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
            cur = nxt;
        }
    CGF_TILE_LOOP_END
    block_sum_to_partials<NV>(v, partials, blockIdx.x);
}

// ============================================================================= Jacobians + PR sums
// raw row data (loaded two rows ahead) and the same row with its two link variables evaluated
template <typename R> struct RawG { typename V2<R>::type p; R pha, phb, eps; unsigned f; };
template <typename R> struct RowG { typename V2<R>::type p; R sa, ca, sb, cb, eps; unsigned f; };

// curl-curl stencils with the boundary doubling of quirk Q10 (cg.h:176-217, 240-282): direct loads
// (they hit L1: the rows were just read for the link phases)
template <typename R>
__device__ __forceinline__ R cgf_curl_a(const Geo &g, int j, size_t n, R H, const R *ae, const R *be, const R *a, const R *b) {
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
__device__ __forceinline__ R cgf_curl_b(const Geo &g, int i, size_t n, R H, const R *ae, const R *be, const R *a, const R *b) {
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

template <typename R, bool SOLVEA, bool PREV, bool EXT>
__global__ void __launch_bounds__(CGF_THREADS, 2)
k_cgf_grad(CgfState<R> S, typename V2<R>::type *__restrict__ gpsi, R *__restrict__ ga, R *__restrict__ gb,
           const typename V2<R>::type *__restrict__ ppsi, const R *__restrict__ pa, const R *__restrict__ pb,
           double *partials) {
    typedef typename V2<R>::type C;
    const Geo &g = S.g;
    const R *const ae_ = EXT ? S.ae : nullptr, *const be_ = EXT ? S.be : nullptr;   // compile-time absent without an external potential
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    // psi and the phases of the two links owned by node (ii, y)
    auto loadraw = [&](int ii, int y) {
        RawG<R> r;
        r.p.x = 0; r.p.y = 0; r.pha = 0; r.phb = 0; r.eps = S.eps; r.f = 0;
        if (ii >= 0 && ii < g.Nx) {
            const size_t n = g.at(ii, y);
            r.f = S.nf[n];
            r.p = S.psi[n];
            R pa_ = 0, pb_ = 0;
            if (ae_) { pa_ += ae_[n]; pb_ += be_[n]; }
            if (S.a) { pa_ += S.a[n]; pb_ += S.b[n]; }
            r.pha = dx * pa_; r.phb = dy * pb_;
            if (S.epsf) r.eps = S.epsf[n];
        }
        return r;
    };
    // exp(-i dx A_a), exp(-i dy A_b) as (sin, cos), only where the link carries weight
    auto link = [&](const RawG<R> &w) {
        RowG<R> r;
        r.p = w.p; r.eps = w.eps; r.f = w.f;
        r.sa = 0; r.ca = 1; r.sb = 0; r.cb = 1;
        if (w.f & (NF_PM | NF_PP)) sincos_r<R>(w.pha, &r.sa, &r.ca);
        if (w.f & (NF_MP | NF_PP)) sincos_r<R>(w.phb, &r.sb, &r.cb);
        return r;
    };
    CGF_TILE_LOOP_BEGIN(30, 1)
        CGF_PF_BEGIN
            CGF_PF(S.psi, sizeof(C)) CGF_PF(S.a, sizeof(R)) CGF_PF(S.b, sizeof(R)) CGF_PF(ae_, sizeof(R))
            CGF_PF(be_, sizeof(R)) CGF_PF(S.epsf, sizeof(R)) CGF_PF(S.nf, 1)
            if (PREV) { CGF_PF(ppsi, sizeof(C)) CGF_PF(pa, sizeof(R)) CGF_PF(pb, sizeof(R)) }
        CGF_PF_END
        RowG<R> prv = link(loadraw(i, ys - 1));
        RowG<R> cur = link(loadraw(i, ys));
        RawG<R> rawn = loadraw(i, ys + 1);
        for (int y = ys; y < ye; y++) {
            const RawG<R> raw2 = loadraw(y + 2 <= ye ? i : -1, y + 2);
            // previous gradient of this node: issued now, used at the end of the iteration
            C q; q.x = 0; q.y = 0;
            R qa = 0, qb = 0;
            if (PREV && in) {
                const size_t n = g.at(i, y);
                q = ppsi[n];
                if (SOLVEA) { qa = pa[n]; qb = pb[n]; }
            }
            const RowG<R> nxt = link(rawn);
            const C pE = shfl_down_c<C>(cur.p), pW = shfl_up_c<C>(cur.p);
            const R sW = __shfl_up_sync(FULL, cur.sa, 1), cW = __shfl_up_sync(FULL, cur.ca, 1);
            if (in) {
                const size_t n = g.at(i, y);
                const unsigned f = cur.f;
                const R eps = cur.eps;
                const C p0 = cur.p;
                C gj;
                gj.x = 0; gj.y = 0;
                if (f) {
                    R wW, wE, wS, wN, gw;
                    du_w<R>(f, wW, wE, wS, wN, gw);
                    const R p = p0.x * p0.x + p0.y * p0.y - eps;
                    gj.x += (R)2.0 * gw * p * p0.x;
                    gj.y += (R)2.0 * gw * p * p0.y;
                    // g_grad_jac_psi(psi0, ph, psi1) = 2 (psi0 - psi1 U(ph))  (cg.h:5-12); the W and S links
                    // enter with the opposite phase: sincos(-x) = (-sin x, cos x)
                    if (f & (NF_MM | NF_MP)) {
                        const C z = gradc<R, C>(p0, -sW, cW, pW);
                        gj.x += wW * idx2 * ((R)-2.0 * z.x); gj.y += wW * idx2 * ((R)-2.0 * z.y);
                    }
                    if (f & (NF_PM | NF_PP)) {
                        const C z = gradc<R, C>(p0, cur.sa, cur.ca, pE);
                        gj.x += wE * idx2 * ((R)-2.0 * z.x); gj.y += wE * idx2 * ((R)-2.0 * z.y);
                    }
                    if (f & (NF_MM | NF_PM)) {
                        const C z = gradc<R, C>(p0, -prv.sb, prv.cb, prv.p);
                        gj.x += wS * idy2 * ((R)-2.0 * z.x); gj.y += wS * idy2 * ((R)-2.0 * z.y);
                    }
                    if (f & (NF_MP | NF_PP)) {
                        const C z = gradc<R, C>(p0, cur.sb, cur.cb, nxt.p);
                        gj.x += wN * idy2 * ((R)-2.0 * z.x); gj.y += wN * idy2 * ((R)-2.0 * z.y);
                    }
                }
                const R dxdy = dx * dy;
                gj.x *= dxdy; gj.y *= dxdy;
                gpsi[n] = gj;
                const bool own = y >= g.j0 && y < g.j1;                     // halo rows (slabs): gradient only, no sums
                if (PREV && own) {
                    v[0] += (double)(gj.x * (gj.x - q.x) + gj.y * (gj.y - q.y));
                    v[1] += (double)(q.x * q.x + q.y * q.y);
                }
                if (SOLVEA) {
                    const R pm = (f & NF_PM) ? (R)1 : (R)0, pp = (f & NF_PP) ? (R)1 : (R)0, mp = (f & NF_MP) ? (R)1 : (R)0;
                    if (i < g.Nx - 1) {
                        R w = S.kappa2 * cgf_curl_a<R>(g, y, n, S.H, ae_, be_, S.a, S.b);
                        if (f & (NF_PM | NF_PP)) {
                            const R js = (p0.x * pE.y - p0.y * pE.x) * cur.ca - (p0.x * pE.x + p0.y * pE.y) * cur.sa;
                            w += -((R)0.5 * (pm + pp)) * idx * js;
                        }
                        w = (R)2.0 * dx * dy * w;
                        ga[n] = w;
                        if (PREV && own) { v[2] += (double)(w * (w - qa)); v[3] += (double)(qa * qa); }
                    }
                    if (y < g.Ny - 1) {
                        R w = S.kappa2 * cgf_curl_b<R>(g, i, n, S.H, ae_, be_, S.a, S.b);
                        if (f & (NF_MP | NF_PP)) {
                            const R js = (p0.x * nxt.p.y - p0.y * nxt.p.x) * cur.cb - (p0.x * nxt.p.x + p0.y * nxt.p.y) * cur.sb;
                            w += -((R)0.5 * (mp + pp)) * idy * js;
                        }
                        w = (R)2.0 * dx * dy * w;
                        gb[n] = w;
                        if (PREV && own) { v[2] += (double)(w * (w - qb)); v[3] += (double)(qb * qb); }
                    }
                }
            }
            prv = cur;
            cur = nxt;
            rawn = raw2;
        }
    CGF_TILE_LOOP_END
    if (PREV) block_sum_to_partials<4>(v, partials, blockIdx.x);
}

// beta = max(num/den, 0) in real_t, nan -> 0 (divide_scalars_positive, utils.h:140-146); stays on the device
template <typename R>
__global__ void k_cgf_beta(const double *__restrict__ sums, double *beta) {
    if (threadIdx.x < 2) {
        R q = (R)sums[2 * threadIdx.x] / (R)sums[2 * threadIdx.x + 1];
        beta[threadIdx.x] = (q > (R)0) ? (double)q : 0.0;
    }
}

// ----------------------------------------------------------------------------- host side
// rows per strip: 32 where that still leaves a few tiles per resident CTA, fewer on small grids
static int cgf_rows(const svl_ctx *c) {
    const Geo &g = c->g;
    for (int V = 32; V > 4; V >>= 1) {
        long tiles = (long)((g.Nx + 30 * CGF_WARPS - 1) / (30 * CGF_WARPS)) * ((g.j1 - g.j0 + V - 1) / V);
        if (tiles >= 148 * 2 * 2) return V;
    }
    return 4;
}
static int cgf_grid(svl_ctx *c, int warps = CGF_WARPS) {
    const Geo &g = c->g;
    const int V = cgf_rows(c);
    int ntiles = ((g.Nx + 30 * warps - 1) / (30 * warps)) * ((g.j1 - g.j0 + V - 1) / V);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
    int cap = nsm * 2 * 4;                       // a few waves of resident CTAs: balances the tail, ~1200 partials
    return ntiles < cap ? ntiles : cap;
}

template <typename R>
static CgfState<R> cgf_state(svl_ctx *c, double kappa2, double eps, const svl_buf *epsf, double H, const svl_buf *psi,
                             const svl_buf *abei, const svl_buf *ab) {
    typedef typename V2<R>::type C;
    CgfState<R> S;
    S.g = c->g;
    S.V = cgf_rows(c);
    // slabs: gradients, directions and updates are also evaluated on two halo rows per neighbour, so only psi / A
    // need an exchange per iteration (their halos are 8 rows deep); see cgf_end_t
    S.ext_lo = (c->slab_on && c->has_lo) ? 2 : 0;
    S.ext_hi = (c->slab_on && c->has_hi) ? 2 : 0;
    S.kappa2 = (R)kappa2; S.eps = (R)eps; S.H = (R)H;
    S.epsf = epsf ? (const R *)epsf->p[0] : nullptr;
    S.nf = c->nf;
    S.psi = (const C *)psi->p[0];
    S.ae = abei ? (const R *)abei->p[0] : nullptr; S.be = abei ? (const R *)abei->p[1] : nullptr;
    S.a = ab ? (const R *)ab->p[0] : nullptr; S.b = ab ? (const R *)ab->p[1] : nullptr;
    return S;
}

template <typename R>
static int cgf_begin_t(svl_ctx *c, int solveA, int have_prev, double kappa2, double eps, const svl_buf *epsf, double H,
                       const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi, svl_buf *g_psi_prev,
                       svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A, double *beta, double *c_out) {
    typedef typename V2<R>::type C;
    const int nb = cgf_grid(c);
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * 17));
    CgfState<R> S = cgf_state<R>(c, kappa2, eps, epsf, H, psi, abei, ab);
    R *gA0 = solveA ? (R *)g_A->p[0] : nullptr, *gA1 = solveA ? (R *)g_A->p[1] : nullptr;
    const R *pA0 = solveA ? (const R *)g_A_prev->p[0] : nullptr, *pA1 = solveA ? (const R *)g_A_prev->p[1] : nullptr;
    const bool ext = abei != nullptr;
#define GRAD_ARGS S, (C *)g_psi->p[0], gA0, gA1, (const C *)g_psi_prev->p[0], pA0, pA1, c->partials
#define GRAD_LAUNCH(SA, PV)                                                                               \
    do {                                                                                                  \
        if (ext) k_cgf_grad<R, SA, PV, true><<<nb, CGF_THREADS, 0, c->stream>>>(GRAD_ARGS);               \
        else k_cgf_grad<R, SA, PV, false><<<nb, CGF_THREADS, 0, c->stream>>>(GRAD_ARGS);                  \
    } while (0)
    if (solveA && have_prev) GRAD_LAUNCH(true, true);
    else if (solveA) GRAD_LAUNCH(true, false);
    else if (have_prev) GRAD_LAUNCH(false, true);
    else GRAD_LAUNCH(false, false);
#undef GRAD_LAUNCH
#undef GRAD_ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    double *dbeta = c->d_result + 32;                 // device-resident beta[2]
    if (have_prev) {
        SVL_TRY(svl_finish_sum(c, nb, 4, 1.0, nullptr));          // d_result[0..3], no host read
        if (c->slab_on) SVL_TRY(svl_board_allsum(c, c->d_result, 4));
        k_cgf_beta<R><<<1, 32, 0, c->stream>>>(c->d_result, dbeta);
        SVL_CHECK(cudaGetLastError());
        c->stat_launches += 1;
    } else {
        // first iteration of a cg() call: keep the betas of the previous call (quirk Q6)
        c->h_result[40] = beta[0]; c->h_result[41] = beta[1];
        SVL_CHECK(cudaMemcpyAsync(dbeta, c->h_result + 40, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    // direction update fused with the coefficients; new directions go to scratch planes, then swap
    svl_buf *dn_psi = nullptr, *dn_A = nullptr;
    if (c->slab_on) {      // the TD scratch planes belong to the exchanged arena and must stay psi / A planes
        if (!c->cg_s_node) SVL_TRY(svl_alloc(c, SVL_NODE_C, 0, 0, &c->cg_s_node));
        if (solveA && !c->cg_s_edge) SVL_TRY(svl_alloc(c, SVL_EDGE, 0, 0, &c->cg_s_edge));
        dn_psi = c->cg_s_node; dn_A = c->cg_s_edge;
    } else {
        SVL_TRY(svl_scratch_node(c, 0, &dn_psi));
        if (solveA) SVL_TRY(svl_scratch_edge(c, 0, &dn_A));
    }
    // Quirk Q11: the coefficient kernels use the scalar eps (0.0 when eps is a field)
    S.eps = (R)(epsf ? 0.0 : eps);
#define COEF17_ARGS S, dbeta, (const C *)g_psi->p[0], gA0, gA1, (const C *)d_psi->p[0], (const R *)d_A->p[0],  \
                    (const R *)d_A->p[1], (C *)dn_psi->p[0], (R *)dn_A->p[0], (R *)dn_A->p[1], c->partials
#define COEF5_ARGS S, dbeta, (const C *)g_psi->p[0], nullptr, nullptr, (const C *)d_psi->p[0], nullptr, nullptr,     \
                   (C *)dn_psi->p[0], nullptr, nullptr, c->partials
    const int nt17 = sizeof(R) == 8 ? 128 : CGF_THREADS, nb17 = sizeof(R) == 8 ? cgf_grid(c, 4) * 3 / 2 : nb;
    if (solveA) SVL_TRY(svl_ensure_partials(c, (size_t)nb17 * 17));
    if (solveA && ext) k_cgf_coef<R, 17, true><<<nb17, nt17, 0, c->stream>>>(COEF17_ARGS);
    else if (solveA) k_cgf_coef<R, 17, false><<<nb17, nt17, 0, c->stream>>>(COEF17_ARGS);
    else if (ext) k_cgf_coef<R, 5, true><<<nb, CGF_THREADS, 0, c->stream>>>(COEF5_ARGS);
    else k_cgf_coef<R, 5, false><<<nb, CGF_THREADS, 0, c->stream>>>(COEF5_ARGS);
#undef COEF17_ARGS
#undef COEF5_ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    SVL_TRY(svl_swap(c, d_psi, dn_psi));
    if (solveA) SVL_TRY(svl_swap(c, d_A, dn_A));
    SVL_CHECK(cudaMemcpyAsync(c->h_result + 32, dbeta, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    const int nvc = solveA ? 17 : 5;
    if (c->slab_on) {      // row slabs: rank-ordered sum over the residual board, then one host sync
        SVL_TRY(svl_finish_sum(c, solveA ? nb17 : nb, nvc, (double)((R)c->g.dx * (R)c->g.dy), nullptr));
        SVL_TRY(svl_board_allsum(c, c->d_result, nvc));
        SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, nvc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < nvc; k++) c_out[k] = c->h_result[k];
    } else {
        SVL_TRY(svl_finish_sum(c, solveA ? nb17 : nb, nvc, (double)((R)c->g.dx * (R)c->g.dy), c_out));   // one host sync
    }
    beta[0] = c->h_result[32];
    if (solveA) beta[1] = c->h_result[33];
    return 0;
}

template <typename R>
static int cgf_end_t(svl_ctx *c, int solveA, double kappa2, double eps, const svl_buf *epsf, double H, svl_buf *psi,
                     const svl_buf *abei, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi,
                     double alpha_A, double *E_out) {
    typedef typename V2<R>::type C;
    const int nb = cgf_grid(c);
    SVL_TRY(svl_ensure_partials(c, (size_t)nb));
    CgfState<R> S = cgf_state<R>(c, kappa2, eps, epsf, H, psi, abei, ab);
    svl_buf *pn = nullptr, *An = nullptr;
    SVL_TRY(svl_scratch_node(c, 0, &pn));
    if (solveA) SVL_TRY(svl_scratch_edge(c, 0, &An));
    const bool ext = abei != nullptr;
#define UPD_A_ARGS S, (const C *)d_psi->p[0], (const R *)d_A->p[0], (const R *)d_A->p[1], (R)alpha_psi, (R)alpha_A,  \
                   (C *)pn->p[0], (R *)An->p[0], (R *)An->p[1], c->partials
#define UPD_P_ARGS S, (const C *)d_psi->p[0], nullptr, nullptr, (R)alpha_psi, (R)0, (C *)pn->p[0], nullptr, nullptr, \
                   c->partials
    if (solveA && ext) k_cgf_update<R, true, true><<<nb, CGF_THREADS, 0, c->stream>>>(UPD_A_ARGS);
    else if (solveA) k_cgf_update<R, true, false><<<nb, CGF_THREADS, 0, c->stream>>>(UPD_A_ARGS);
    else if (ext) k_cgf_update<R, false, true><<<nb, CGF_THREADS, 0, c->stream>>>(UPD_P_ARGS);
    else k_cgf_update<R, false, false><<<nb, CGF_THREADS, 0, c->stream>>>(UPD_P_ARGS);
#undef UPD_A_ARGS
#undef UPD_P_ARGS
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    SVL_TRY(svl_swap(c, psi, pn));
    if (solveA) SVL_TRY(svl_swap(c, ab, An));
    double E = 0.0;
    if (c->slab_on) {      // row slabs: refresh the 8-row halos of the new state, add the energies
        SVL_TRY(svl_slab_push_psi(c, psi));
        if (solveA) SVL_TRY(svl_slab_push_ab(c, ab));
        SVL_TRY(svl_slab_wait(c));
        SVL_TRY(svl_finish_sum(c, nb, 1, (double)((R)c->g.dx * (R)c->g.dy), nullptr));
        SVL_TRY(svl_board_allsum(c, c->d_result, 1));
        SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        E = c->h_result[0];
    } else
    SVL_TRY(svl_finish_sum(c, nb, 1, (double)((R)c->g.dx * (R)c->g.dy), &E));
    if (E_out) *E_out = E;
    return 0;
}

int svl_cgf_begin(svl_ctx *c, int solveA, int have_prev, double kappa2, double eps, const svl_buf *epsf, double H,
                  const svl_buf *psi, const svl_buf *abei, const svl_buf *ab, svl_buf *g_psi, svl_buf *g_psi_prev,
                  svl_buf *d_psi, svl_buf *g_A, svl_buf *g_A_prev, svl_buf *d_A, double *beta, double *c_out) {
    if (c->rsize == 4) return cgf_begin_t<float>(c, solveA, have_prev, kappa2, eps, epsf, H, psi, abei, ab, g_psi, g_psi_prev,
                                                 d_psi, g_A, g_A_prev, d_A, beta, c_out);
    return cgf_begin_t<double>(c, solveA, have_prev, kappa2, eps, epsf, H, psi, abei, ab, g_psi, g_psi_prev, d_psi, g_A,
                               g_A_prev, d_A, beta, c_out);
}

int svl_cgf_end(svl_ctx *c, int solveA, double kappa2, double eps, const svl_buf *epsf, double H, svl_buf *psi,
                const svl_buf *abei, svl_buf *ab, const svl_buf *d_psi, const svl_buf *d_A, double alpha_psi,
                double alpha_A, double *E_out) {
    if (c->rsize == 4) return cgf_end_t<float>(c, solveA, kappa2, eps, epsf, H, psi, abei, ab, d_psi, d_A, alpha_psi, alpha_A, E_out);
    return cgf_end_t<double>(c, solveA, kappa2, eps, epsf, H, psi, abei, ab, d_psi, d_A, alpha_psi, alpha_A, E_out);
}
