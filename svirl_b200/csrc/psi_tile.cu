// Temporally blocked psi sweeps, register-resident variant: K sweeps per launch on an
// overlapped 2-D tile (same arithmetic as k_psi_sweep in td.cu / svirl/cuda/td.h:5-133).
//
//   * a CTA loads an extended tile of TXE x EY nodes (interior + halo of H = roundup(K,4)
//     columns / K rows) with one set of 2-D TMA boxes (psi, rhs, a, b, [eps], flags) into shared
//     memory; zero fill outside the plane supplies the domain boundary;
//   * everything that is constant during the solve -- right-hand side, 1/diagonal and the two link
//     coefficients w*dt/d^2*exp(-i d A) of a thread's nodes (one sincos per link and launch; fp32: MUFU after an
//     exact reduction, link_sincos) -- lives in REGISTERS; the material flags of a node index a 16-entry shared
//     table of link weights / neighbour term / activity; only the psi iterate goes through shared memory (double
//     buffered, one __syncthreads per sweep);
//   * k_psi_tile: a thread owns V consecutive rows of ONE COLUMN (26 shared loads + 8 stores per sweep for 8
//     nodes); k_psi_patch (fp32, everything but the boundary tile rows of a slab): a thread owns a PATCH of
//     2 columns x 4 rows with the iterate in separate planes for the even and the odd columns (16 + 8).  Both do
//     the same 16 chained FMAs per node in the same order: bit-identical results;
//   * the halo shrinks by one ring per sweep; after K sweeps the interior is exact and is
//     written back; the max-norm update of each of the K sweeps is reduced over the interior
//     (REDUX per warp, shared atomicMax) and merged with one global atomicMax per CTA and sweep;
//   * launches of one solve are chained as programmatic dependent launches (griddepcontrol); a launch issued ahead
//     of the host's stop decision carries a gate word (td.cu: pipelined solves); on row slabs the boundary tile
//     rows run in a launch of their own on a few CTAs (halo-flag wait, peer stores, one fence per CTA).
// HBM traffic per launch: the one-sweep bytes times the halo overhead (1.3-1.5x), for K sweeps.
#include "common.cuh"
#include <cuda.h>

struct TileArgs {
    Geo g;
    double dt, eps, lang_c;
    uint32_t rand_t;
    int noise;
    int same_rhs;
    void *out;
    unsigned long long *slots;
    // slabs: halo rows of the input must have arrived (epoch flags written by the neighbours)
    const unsigned long long *wait_flags;
    unsigned long long wait_epoch;
    int has_lo, has_hi;
    SlabPush push;                 // slabs: in-kernel push of the output's boundary rows
    int push_expect[2];            // tiles that contribute to the lo / hi push
    int row0, nrow, row1, nrow1;   // tile rows this launch covers: [row0, row0+nrow) then [row1, row1+nrow1)
    int permute;                   // slabs, single launch: process the first / last tile row last
    int pdl_trigger;               // release a programmatic dependent launch at once (slab boundary launch -> interior launch;
                                   // one GPU: batch n -> batch n+1, whose CTAs then start up in this launch's tail)
    int pdl_wait;                  // this launch may have started before its predecessor ended: wait before the first load
                                   // (pdl_trigger == 2: release the dependent launch only after that wait)
    int pdl_wait_end;              // slabs, interior launch: do not COMPLETE before the boundary launch it ran beside has
    const int *gate;               // pre-issued launch (td.cu, pipelined solves): do nothing unless *gate != 0
    int defer_publish;             // slabs, boundary launch: a CTA fences and reports its pushed tiles ONCE, after its last tile
    SpinGuard sg;                  // bound of the spin waits (halo flags, TMA barrier)
    unsigned long long *trace;     // slabs, diagnostics: [0] first CTA start, [1] last CTA end, [2] longest flag wait,
                                   // [3] time the last flag wait ended (all %globaltimer ns), or null
};

__device__ __forceinline__ uint32_t t_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(t_smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(t_smem_u32(bar))
        : "memory");
}

template <typename R, int K, int TXE, int V, int NB, bool EPS>
struct TileSmem {
    typedef typename V2<R>::type C;
    static constexpr int EY = V * NB;
    static constexpr int H = ((K + 3) / 4) * 4;
    static constexpr int NFW = TXE + 16;
    static constexpr int XW = TXE + 2;                  // exchange row width (zero pad column each side)
    static constexpr int XR = EY + 2;                   // exchange rows (zero pad row each side)
    // staging (TMA destinations, 128-byte aligned); refilled for the NEXT tile while this one is swept
    static constexpr size_t st_psi = 0;
    static constexpr size_t st_rhs = st_psi + sizeof(C) * EY * TXE;
    static constexpr size_t st_a = st_rhs + sizeof(C) * EY * TXE;
    static constexpr size_t st_b = st_a + sizeof(R) * EY * TXE;
    static constexpr size_t st_eps = st_b + sizeof(R) * EY * TXE;
    static constexpr size_t st_nf = st_eps + (EPS ? sizeof(R) * EY * TXE : 0);
    static constexpr size_t st_end = ((st_nf + (size_t)EY * NFW + 127) / 128) * 128;
    static constexpr uint32_t tx_bytes = (uint32_t)(2 * sizeof(C) * EY * TXE + (EPS ? 3 : 2) * sizeof(R) * EY * TXE + EY * NFW);
    // psi exchange (double buffered) and the a-link coefficient tile (W coefficient of the right neighbour)
    static constexpr size_t off_x0 = st_end;
    static constexpr size_t off_x1 = off_x0 + sizeof(C) * XR * XW;
    static constexpr size_t off_la = off_x1 + sizeof(C) * XR * XW;
    static constexpr size_t off_bar = ((off_la + sizeof(C) * XR * XW + 15) / 16) * 16;
    static constexpr size_t total = off_bar + 16;
};

// Persistent CTAs: each loops over tiles tile = blockIdx.x, +gridDim.x, ...; the TMA boxes of the
// next tile are issued as soon as this tile's constants are in registers, so the load overlaps the
// K sweeps.
// Link variable cos/sin of the tile kernel.  LINKS = 0: the library's polynomial sincos (<= 2 ulp).
// LINKS = 1 (fp32 only, option psi_links): two-term Cody-Waite reduction by 2 pi + the hardware's
// sin.approx / cos.approx (SASS MUFU.SIN / MUFU.COS): 7 instructions instead of 29 per link, max abs
// error ~6e-7 -- the size of the rounding error of the fp32 phase d*A itself once |d*A| > 8.
template <typename R, int LINKS>
__device__ __forceinline__ void link_sincos(R x, R *s, R *c) { sincos_r<R>(x, s, c); }
template <>
__device__ __forceinline__ void link_sincos<float, 1>(float x, float *s, float *c) {
    // no large-argument branch: one FMA rounding for any |x| < 2^23, where the fp32 phase itself is already
    // uncertain by more than a radian
    const float j = rintf(x * 0.15915494309189535f);
    float r = fmaf(j, -6.2831854820251465f, x);          // 2 pi = 6.2831854820251465 - 1.7484556000744883e-07
    r = fmaf(j, 1.7484556000744883e-07f, r);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(*s) : "f"(r));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(*c) : "f"(r));
}
// The two links of a node together: both polynomial cores run unconditionally and side by side (their constants are
// fetched once, the two Horner chains interleave) and ONE rarely taken test covers both arguments.
template <typename R, int LINKS>
__device__ __forceinline__ void link_sincos2(R xa, R xb, R *sa, R *ca, R *sb, R *cb) {
    sincos_fast(xa, sa, ca);
    sincos_fast(xb, sb, cb);
    if (!(sincos_fast_ok(xa) && sincos_fast_ok(xb))) { sincos_any(xa, sa, ca); sincos_any(xb, sb, cb); }
}
template <>
__device__ __forceinline__ void link_sincos2<float, 1>(float xa, float xb, float *sa, float *ca, float *sb, float *cb) {
    link_sincos<float, 1>(xa, sa, ca);
    link_sincos<float, 1>(xb, sb, cb);
}
// 1/D of the Jacobi diagonal (D of order 1, never denormal): hardware seed + Newton steps, <= 1 ulp, no slow path
__device__ __forceinline__ float rcp_diag(float x) { return rcp_r(x); }
__device__ __forceinline__ double rcp_diag(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));         // MUFU.RCP64H: ~20 bits
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    const double e = fma(-x, r, 1.0);                              // third step: the residual of a 40+ bit estimate
    return fma(r, e, r);
}

template <typename R, int K, int TXE, int V, int NB, bool EPS, bool SLAB, int LINKS>
__global__ void __launch_bounds__(TXE *NB, ((sizeof(R) == 4 && V * NB <= 32) ? 2 : 1))
k_psi_tile(const __grid_constant__ TileArgs A, const __grid_constant__ CUtensorMap tm_psi,
           const __grid_constant__ CUtensorMap tm_rhs, const __grid_constant__ CUtensorMap tm_a,
           const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_eps,
           const __grid_constant__ CUtensorMap tm_nf) {
    typedef typename V2<R>::type C;
    typedef TileSmem<R, K, TXE, V, NB, EPS> S;
    constexpr int EY = S::EY, H = S::H, NFW = S::NFW, XW = S::XW, XR = S::XR;
    constexpr int TX = TXE - 2 * H, TYO = EY - 2 * K, NT = TXE * NB;
    extern __shared__ __align__(128) unsigned char smem[];
    C *xb0 = (C *)(smem + S::off_x0);
    C *xb1 = (C *)(smem + S::off_x1);
    C *sla = (C *)(smem + S::off_la);
    uint64_t *bar = (uint64_t *)(smem + S::off_bar);

    // A launch issued ahead of the host's decision runs only if the device-side stop rule agreed (td.cu: k_psi_gate)
    if (A.gate && *(const volatile int *)A.gate == 0) return;
    const Geo &g = A.g;
    const int tid = threadIdx.x;
    const int col = tid % TXE, band = tid / TXE;
    const int r0 = band * V;                             // first tile row of this thread
    const int ntx = (g.Nx + TX - 1) / TX, nty = (g.j1 - g.j0 + TYO - 1) / TYO;
    const int ntiles = (SLAB && A.permute) ? ntx * nty : ntx * (A.nrow + A.nrow1);
    const R dt = (R)A.dt, dx = (R)g.dx, dy = (R)g.dy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    const R cx = dt * idx2, cy = dt * idy2, eps0 = (R)A.eps, lang = (R)A.lang_c;

    // Tile rows of this launch: [row0, row0+nrow) then [row1, row1+nrow1).  Slabs, single launch
    // (permute): those two ranges are the boundary rows (they read a neighbour's halo rows and feed the
    // push) and are processed FIRST -- the neighbour's previous push landed a whole launch ago, so
    // the flag wait is free, and this launch's push leaves early -- followed by the interior rows.
    auto tile_of = [&](int q) {
        const int r = q / ntx, nb = A.nrow + A.nrow1;
        int by = r < A.nrow ? A.row0 + r : A.row1 + (r - A.nrow);
        if (SLAB && A.permute && r >= nb) by = A.nrow + (r - nb);        // interior rows [nrow, nty - nrow1)
        return by * ntx + q % ntx;
    };
    bool flags_seen[2] = {false, false};                 // thread 0 only
    int pend[2] = {0, 0};                                // slabs: tiles of this CTA whose push is not reported yet
    auto gtime = [&]() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    auto wait_side = [&](int sdir) {
        if (flags_seen[sdir] || !(sdir == 0 ? A.has_lo : A.has_hi)) return;
        const unsigned long long g0 = A.trace ? gtime() : 0ull;
        svl_spin_ge(A.wait_flags + sdir, A.wait_epoch, A.sg);
        if (A.trace) { const unsigned long long g1 = gtime(); atomicMax(A.trace + 2, g1 - g0); atomicMax(A.trace + 3, g1); }
        flags_seen[sdir] = true;
    };
    auto issue = [&](int q) {                            // thread 0 only
        const int tile = tile_of(q);
        const int bx = tile % ntx, by = tile / ntx;
        if (SLAB) {      // tile rows whose input box reaches into a neighbour's halo rows (K rows beyond the tile)
            const int rows_own = g.j1 - g.j0, top = (by + 1) * TYO < rows_own ? (by + 1) * TYO : rows_own;
            if (by * TYO - K < 0) wait_side(0);
            if (top + K > rows_own) wait_side(1);
        }
        const int xg0 = bx * TX - H, prow = g.j0 + by * TYO - K - g.rb;
        const int xs16 = ((xg0 + 1024) / 16) * 16 - 1024;
        uint32_t bytes = S::tx_bytes;
        if (A.same_rhs) bytes -= (uint32_t)(sizeof(C) * EY * TXE);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem_u32(bar)), "r"(bytes) : "memory");
        const int cmul = sizeof(C) / 8 == 2 ? 2 : 1;
        t_tma_load_2d(smem + S::st_psi, &tm_psi, xg0 * cmul, prow, bar);
        if (!A.same_rhs) t_tma_load_2d(smem + S::st_rhs, &tm_rhs, xg0 * cmul, prow, bar);
        t_tma_load_2d(smem + S::st_a, &tm_a, xg0, prow, bar);
        t_tma_load_2d(smem + S::st_b, &tm_b, xg0, prow, bar);
        if (EPS) t_tma_load_2d(smem + S::st_eps, &tm_eps, xg0, prow, bar);
        t_tma_load_2d(smem + S::st_nf, &tm_nf, xs16, prow, bar);
    };

    int tile = blockIdx.x;
    if (A.pdl_trigger == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (SLAB && A.trace && tid == 0) atomicMin(A.trace + 0, gtime());
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // everything this launch reads or overwrites comes after its first TMA load, which waits for the predecessor
        if (A.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (tile < ntiles) issue(tile);
    }
    // zero the pad ring of the exchange / coefficient tiles once (never overwritten afterwards)
    {
        C z; z.x = 0; z.y = 0;
        for (int i = tid; i < XW; i += NT) {
            xb0[i] = z; xb1[i] = z; sla[i] = z;
            xb0[(XR - 1) * XW + i] = z; xb1[(XR - 1) * XW + i] = z; sla[(XR - 1) * XW + i] = z;
        }
        for (int i = tid; i < XR; i += NT) {
            xb0[i * XW] = z; xb1[i * XW] = z; sla[i * XW] = z;
            xb0[i * XW + XW - 1] = z; xb1[i * XW + XW - 1] = z; sla[i * XW + XW - 1] = z;
        }
    }
    if (A.pdl_trigger == 2) {                // chained slab batches: the interior launch may go once the previous batch is over
        __syncthreads();                     // thread 0 is past its griddepcontrol.wait
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    __shared__ unsigned int sm_rmax[K];      // per-sweep max-norm update of this CTA (bit patterns of floats/doubles >= 0)
    __shared__ unsigned long long sm_rmax64[K];
    if (tid < K) { sm_rmax[tid] = 0u; sm_rmax64[tid] = 0ull; }
    // what the four "cell is material" bits of a node mean for its update, tabulated once per CTA: E / N link
    // weights (0 or dt/d^2), the neighbour-count term of the diagonal, node active (1 or 0); td.h:48-116
    __shared__ __align__(16) R s_fl[16][4];
    if (tid < 16) {
        const unsigned f = tid;
        const R nwx = ((f & (NF_MM | NF_MP)) ? (R)1 : (R)0) + ((f & (NF_PM | NF_PP)) ? (R)1 : (R)0);
        const R nwy = ((f & (NF_MM | NF_PM)) ? (R)1 : (R)0) + ((f & (NF_MP | NF_PP)) ? (R)1 : (R)0);
        s_fl[f][0] = (f & (NF_PM | NF_PP)) ? cx : (R)0;
        s_fl[f][1] = (f & (NF_MP | NF_PP)) ? cy : (R)0;
        s_fl[f][2] = idx2 * nwx + idy2 * nwy;
        s_fl[f][3] = f ? (R)1 : (R)0;
    }
    uint32_t phase = 0;

    for (; tile < ntiles; tile += gridDim.x) {           // `tile` counts positions in the processing order
        const int tcur = tile_of(tile);
        const int bx = tcur % ntx, by = tcur / ntx;
        const int xg0 = bx * TX - H;                     // global column of tile column 0
        const int yg0 = g.j0 + by * TYO - K;             // global row of tile row 0
        const int x = xg0 + col;
        const int nfd = xg0 - (((xg0 + 1024) / 16) * 16 - 1024);
        // ---- wait for this tile's boxes: one warp polls, the barrier releases the rest
        if (tid < 32) {
            // (thread 0 may be spinning on a neighbour's halo flag before it issues the boxes: no iteration
            // bound here; a broken descriptor faults the launch by itself)
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(t_smem_u32(bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        __syncthreads();      // also: everybody is done with the previous tile's exchange buffers

        // ---- Langevin noise (td.h:92-101): folded into the staged right-hand side by a pass of its own, so that
        // the unrolled constants loop below carries no noise code (each thread touches its own nodes only)
        const bool rhs_is_psi = A.same_rhs && !A.noise;
        if (A.noise) {
            const bool xin0 = (x >= 0 && x < g.Nx);
#pragma unroll 1
            for (int v = 0; v < V; v++) {
                const int r = r0 + v, si = r * TXE + col;
                C qq = A.same_rhs ? ((const C *)(smem + S::st_psi))[si] : ((const C *)(smem + S::st_rhs))[si];
                const unsigned f = xin0 ? (smem + S::st_nf)[r * NFW + nfd + col] : 0u;
                if (f) {
                    uint32_t nn = (uint32_t)x + (uint32_t)g.Nx * (uint32_t)(yg0 + r);
                    qq.x += lang * (rand_1<R>(nn, A.rand_t) - (R)0.5);
                    qq.y += lang * (rand_2<R>(nn, A.rand_t) - (R)0.5);
                }
                ((C *)(smem + S::st_rhs))[si] = qq;
            }
            // the next tile's TMA boxes (async proxy) overwrite what this thread just wrote through the generic proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        // ---- per-node constants into registers
        C psi[V], q[V], La[V], Lb[V];
        R di[V];
        C LbS0;
        LbS0.x = 0; LbS0.y = 0;
        const bool xin = (x >= 0 && x < g.Nx);
        if (r0 > 0) {         // S-link coefficient of the strip's first row = b-link of tile row r0-1
            const int si = (r0 - 1) * TXE + col;
            unsigned f = (smem + S::st_nf)[(r0 - 1) * NFW + nfd + col];
            if (!xin) f = 0;
            const R wN = s_fl[f][1];
            R sn, cs;
            link_sincos<R, LINKS>(dy * ((const R *)(smem + S::st_b))[si], &sn, &cs);
            LbS0.x = wN * cs; LbS0.y = wN * sn;
        }
#pragma unroll
        for (int v = 0; v < V; v++) {
            const int r = r0 + v, si = r * TXE + col;
            C p0 = ((const C *)(smem + S::st_psi))[si];
            C qq = rhs_is_psi ? p0 : ((const C *)(smem + S::st_rhs))[si];
            R av = ((const R *)(smem + S::st_a))[si], bv = ((const R *)(smem + S::st_b))[si];
            unsigned f = (smem + S::st_nf)[r * NFW + nfd + col];
            R e = EPS ? ((const R *)(smem + S::st_eps))[si] : eps0;
            if (!xin) f = 0;
            R wE, wN, nw, act;
            if (sizeof(R) == 4) {
                const float4 w = *(const float4 *)&s_fl[f][0];
                wE = w.x; wN = w.y; nw = w.z; act = w.w;
            } else {
                const double2 w0 = *(const double2 *)&s_fl[f][0], w1 = *(const double2 *)&s_fl[f][2];
                wE = w0.x; wN = w0.y; nw = w1.x; act = w1.y;
            }
            R sa, ca, sb, cb;
            C la, lb;
            link_sincos2<R, LINKS>(dx * av, dy * bv, &sa, &ca, &sb, &cb);
            la.x = wE * ca; la.y = wE * sa;
            lb.x = wN * cb; lb.y = wN * sb;
            qq.x *= act; qq.y *= act;
            const R D = (R)1.0 + dt * (qq.x * qq.x + qq.y * qq.y - e + nw);
            // inactive / out-of-domain nodes stay exactly 0 (td.h:117 writes psi_next = 0): D may vanish there
            // (dt*eps == 1), and 0 * inf would seed NaNs that the zero-weight links then spread
            const R d = f ? rcp_diag(D) : (R)0;
            psi[v] = p0; q[v] = qq; La[v] = la; Lb[v] = lb; di[v] = d;
            const int xi = (r + 1) * XW + col + 1;
            xb0[xi] = p0;
            sla[xi] = la;
        }
        __syncthreads();      // staging fully consumed; level-0 values and a-link coefficients visible
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);   // prefetch overlaps the sweeps

        const bool cin = (col >= H && col < TXE - H && x < g.Nx);
        unsigned inmask = 0;                  // bit v: node v of this thread is an output node of the tile
#pragma unroll
        for (int v = 0; v < V; v++)
            if (cin && r0 + v >= K && r0 + v < EY - K && yg0 + r0 + v < g.j1) inmask |= 1u << v;
        // One sweep: iterate `in` (registers; W/E neighbours and the strip ends from `src`) -> `out` (registers, and
        // `dst` for the neighbouring threads unless it is the last sweep).  The k loop below is unrolled by two with
        // the roles of the two register sets swapped, so no sweep ends with a register copy; it is not unrolled
        // further because the body would not fit the instruction cache.
        auto sweep = [&](const C(&in)[V], C(&out)[V], const C *src, C *dst, const int k) {
            const C below = src[(r0) * XW + col + 1];              // tile row r0-1
            const C above = src[(r0 + V + 1) * XW + col + 1];      // tile row r0+V
#pragma unroll
            for (int v = 0; v < V; v++) {
                const int xi = (r0 + v + 1) * XW + col + 1;
                const C pw = src[xi - 1], pe = src[xi + 1];
                const C lw = sla[xi - 1];
                const C pS = v > 0 ? in[v - 1] : below;
                const C pN = v < V - 1 ? in[v + 1] : above;
                const C ls = v > 0 ? Lb[v - 1] : LbS0;
                // 16 chained FMAs: W,S use (c + i s) psi, E,N use (c - i s) psi
                R ax = q[v].x, ay = q[v].y;
                ax = fma_r(lw.x, pw.x, ax);     ay = fma_r(lw.x, pw.y, ay);
                ax = fma_r(-lw.y, pw.y, ax);    ay = fma_r(lw.y, pw.x, ay);
                ax = fma_r(La[v].x, pe.x, ax);  ay = fma_r(La[v].x, pe.y, ay);
                ax = fma_r(La[v].y, pe.y, ax);  ay = fma_r(-La[v].y, pe.x, ay);
                ax = fma_r(ls.x, pS.x, ax);     ay = fma_r(ls.x, pS.y, ay);
                ax = fma_r(-ls.y, pS.y, ax);    ay = fma_r(ls.y, pS.x, ay);
                ax = fma_r(Lb[v].x, pN.x, ax);  ay = fma_r(Lb[v].x, pN.y, ay);
                ax = fma_r(Lb[v].y, pN.y, ax);  ay = fma_r(-Lb[v].y, pN.x, ay);
                out[v].x = ax * di[v];
                out[v].y = ay * di[v];
            }
            const bool more = k < K - 1;
            // max-norm update over this thread's output nodes, one shared atomicMax per warp and sweep.  Non-negative
            // values order like their bit patterns: fp32 reduces with REDUX on the bits; fp64 keeps the running maximum
            // as a 64-bit integer as well (a double fmax is ~9 instructions on this machine, an integer one 4) and
            // reduces the high and low words with two REDUX instead of five shuffle + fmax rounds.
            if (sizeof(R) == 4) {
                R rm = 0;
#pragma unroll
                for (int v = 0; v < V; v++) {
                    if (inmask & (1u << v))
                        rm = fmax(rm, fmax(fabs(out[v].x - in[v].x), fabs(out[v].y - in[v].y)));
                    if (more) dst[(r0 + v + 1) * XW + col + 1] = out[v];
                }
                const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint((float)rm));
                if ((tid & 31) == 0 && wm) atomicMax(&sm_rmax[k], wm);
            } else {
                unsigned long long rb = 0ull;
#pragma unroll
                for (int v = 0; v < V; v++) {
                    const double ex = (double)(out[v].x - in[v].x), ey = (double)(out[v].y - in[v].y);
                    const unsigned long long bx = ((unsigned long long)(__double2hiint(ex) & 0x7fffffff) << 32) | (unsigned)__double2loint(ex);
                    const unsigned long long by = ((unsigned long long)(__double2hiint(ey) & 0x7fffffff) << 32) | (unsigned)__double2loint(ey);
                    const unsigned long long bm = bx > by ? bx : by;
                    rb = ((inmask & (1u << v)) && bm > rb) ? bm : rb;
                    if (more) dst[(r0 + v + 1) * XW + col + 1] = out[v];
                }
                const unsigned hi = (unsigned)(rb >> 32);
                const unsigned whi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned wlo = __reduce_max_sync(0xffffffffu, hi == whi ? (unsigned)rb : 0u);
                if ((tid & 31) == 0 && (whi | wlo)) atomicMax(&sm_rmax64[k], ((unsigned long long)whi << 32) | wlo);
            }
            if (more) __syncthreads();
        };
        {
            C nx[V];
#pragma unroll 1
            for (int k = 0; k + 1 < K; k += 2) {
                sweep(psi, nx, xb0, xb1, k);
                sweep(nx, psi, xb1, xb0, k + 1);
            }
            if (K & 1) {
                sweep(psi, nx, xb0, xb1, K - 1);
#pragma unroll
                for (int v = 0; v < V; v++) psi[v] = nx[v];
            }
        }
        // ---- write the interior (slabs: the first / last `depth` rows also go to the neighbours' halos)
        const int rows_own = g.j1 - g.j0;
        const bool plo = SLAB && A.push.peer[0][0] && by * TYO < A.push.depth;
        const bool phi = SLAB && A.push.peer[1][0] && ((by + 1) * TYO < rows_own ? (by + 1) * TYO : rows_own) > rows_own - A.push.depth;
        {
            // one 64-bit address per thread, then a constant row stride: the stores are predicated, not branched around
            C *o = (C *)A.out + ((long long)(yg0 + r0 - g.rb) * g.P + x);
#pragma unroll
            for (int v = 0; v < V; v++) {
                if (inmask & (1u << v)) o[(long long)v * g.P] = psi[v];
            }
            if (SLAB && (plo || phi)) {
#pragma unroll
                for (int v = 0; v < V; v++) {
                    if (inmask & (1u << v)) {
                        const int y = yg0 + r0 + v;
                        if (plo && y < g.j0 + A.push.depth) ((C *)A.push.peer[0][0])[(size_t)(y - A.push.peer_rb[0]) * g.P + x] = psi[v];
                        if (phi && y >= g.j1 - A.push.depth) ((C *)A.push.peer[1][0])[(size_t)(y - A.push.peer_rb[1]) * g.P + x] = psi[v];
                    }
                }
            }
        }
        if (plo || phi) {                        // CTA-uniform
            if (A.defer_publish) {               // boundary launch: several tiles per CTA, one fence at the end
                pend[0] += plo ? 1 : 0; pend[1] += phi ? 1 : 0;
            } else {
                __syncthreads();                     // all peer stores of the CTA issued ...
                if (tid == 0) {
                    __threadfence_system();          // ... and made visible by ONE fence (grid-sync pattern)
                    for (int sdir = 0; sdir < 2; sdir++) {
                        if (!(sdir == 0 ? plo : phi)) continue;
                        unsigned int done = atomicAdd(&A.push.count[sdir], 1u);
                        if ((int)done == A.push_expect[sdir] - 1) {
                            A.push.count[sdir] = 0;
                            __threadfence_system();
                            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.push.flag[sdir]), "l"(A.push.epoch) : "memory");
                        }
                    }
                }
            }
        }
    }
    if (SLAB && (pend[0] | pend[1])) {           // CTA-uniform
        // The system-scope fence behind peer stores takes ~10-20 us; a boundary CTA therefore pushes all its tiles
        // first and pays it once (the launch gives the boundary tiles to a few CTAs, see launch_tile_t).
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            for (int sdir = 0; sdir < 2; sdir++) {
                if (!pend[sdir]) continue;
                unsigned int done = atomicAdd(&A.push.count[sdir], (unsigned int)pend[sdir]);
                if ((int)done + pend[sdir] == A.push_expect[sdir]) {
                    A.push.count[sdir] = 0;
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.push.flag[sdir]), "l"(A.push.epoch) : "memory");
                }
            }
        }
    }
    if (SLAB && A.trace && tid == 0) atomicMax(A.trace + 1, gtime());
    if (A.pdl_wait_end && tid == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    // ---- per-sweep max-norm updates -> one global atomicMax per CTA and sweep
    __syncthreads();
    if (tid < K) {
        double r = sizeof(R) == 4 ? (double)__uint_as_float(sm_rmax[tid]) : __longlong_as_double((long long)sm_rmax64[tid]);
        if (r > 0.0) atomicMax(A.slots + tid, (unsigned long long)__double_as_longlong(r));
    }
}

// ------------------------------------------------------------------------------- 2 x 4 patch variant (fp32)
// Same tile, same staging, same arithmetic per node (bit-identical results) as k_psi_tile<float, K, 64, 8, 4>, but a
// thread owns a PATCH of 2 columns x 4 rows instead of a column of 8 rows: lane l of warp w holds columns 2l, 2l+1 of
// tile rows 4w .. 4w+3.  The E neighbour of the even column and the W neighbour (and W link coefficient) of the odd
// column are then the thread's own registers, and the iterate is exchanged through separate planes for the even and the
// odd columns, so that what a warp reads is contiguous: 16 LDS.64 + 8 STS.64 per sweep and thread instead of 26 + 8
// (the shared-memory pipe was 55 % busy next to 61 % of the issue slots), LDS.128 / LDS.64 instead of LDS.64 / LDS.32
// in the constants loop.  Non-slab launches only (one GPU, and the interior launch of a slab batch); the boundary tile
// rows of a slab keep the column kernel -- every node gets the same bits from either.
template <int K, bool EPS, int LINKS>
__global__ void __launch_bounds__(256, 2)
k_psi_patch(const __grid_constant__ TileArgs A, const __grid_constant__ CUtensorMap tm_psi,
            const __grid_constant__ CUtensorMap tm_rhs, const __grid_constant__ CUtensorMap tm_a,
            const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_eps,
            const __grid_constant__ CUtensorMap tm_nf) {
    typedef float R;
    typedef float2 C;
    typedef TileSmem<float, K, 64, 8, 4, EPS> S;
    constexpr int TXE = 64, EY = S::EY, H = S::H, NFW = S::NFW, PR = 4, XR = EY + 2, PW = 33;
    constexpr int TX = TXE - 2 * H, TYO = EY - 2 * K, NT = 256;
    static_assert(5 * XR * PW * sizeof(C) <= 3 * sizeof(C) * S::XR * S::XW, "patch planes must fit the exchange area");
    extern __shared__ __align__(128) unsigned char smem[];
    C *pl = (C *)(smem + S::off_x0);
    C *xe[2] = {pl, pl + 2 * XR * PW};                 // even columns: column 2l at index l, index 32 = zero pad (column 64)
    C *xo[2] = {pl + XR * PW, pl + 3 * XR * PW};       // odd columns: column 2l+1 at index l+1, index 0 = zero pad (column -1)
    C *slo = pl + 4 * XR * PW;                         // E-link coefficient of the odd columns (= W coefficient of the even ones)
    uint64_t *bar = (uint64_t *)(smem + S::off_bar);

    if (A.gate && *(const volatile int *)A.gate == 0) return;
    const Geo &g = A.g;
    const int tid = threadIdx.x, lane = tid & 31, r0 = (tid >> 5) * PR, c0 = 2 * lane;
    const int ntx = (g.Nx + TX - 1) / TX;
    const int ntiles = ntx * (A.nrow + A.nrow1);
    const R dt = (R)A.dt, dx = (R)g.dx, dy = (R)g.dy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    const R cx = dt * idx2, cy = dt * idy2, eps0 = (R)A.eps, lang = (R)A.lang_c;
    auto tile_of = [&](int q) {
        const int r = q / ntx;
        const int by = r < A.nrow ? A.row0 + r : A.row1 + (r - A.nrow);
        return by * ntx + q % ntx;
    };
    auto issue = [&](int q) {                            // thread 0 only
        const int tile = tile_of(q);
        const int bx = tile % ntx, by = tile / ntx;
        const int xg0 = bx * TX - H, prow = g.j0 + by * TYO - K - g.rb;
        const int xs16 = ((xg0 + 1024) / 16) * 16 - 1024;
        uint32_t bytes = S::tx_bytes;
        if (A.same_rhs) bytes -= (uint32_t)(sizeof(C) * EY * TXE);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t_smem_u32(bar)), "r"(bytes) : "memory");
        t_tma_load_2d(smem + S::st_psi, &tm_psi, xg0, prow, bar);
        if (!A.same_rhs) t_tma_load_2d(smem + S::st_rhs, &tm_rhs, xg0, prow, bar);
        t_tma_load_2d(smem + S::st_a, &tm_a, xg0, prow, bar);
        t_tma_load_2d(smem + S::st_b, &tm_b, xg0, prow, bar);
        if (EPS) t_tma_load_2d(smem + S::st_eps, &tm_eps, xg0, prow, bar);
        t_tma_load_2d(smem + S::st_nf, &tm_nf, xs16, prow, bar);
    };

    int tile = blockIdx.x;
    if (A.pdl_trigger == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(t_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (A.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (tile < ntiles) issue(tile);
    }
    {   // all five planes to zero once: the pad entries are never written afterwards
        C z; z.x = 0; z.y = 0;
        for (int i = tid; i < 5 * XR * PW; i += NT) pl[i] = z;
    }
    __shared__ unsigned int sm_rmax[K];
    if (tid < K) sm_rmax[tid] = 0u;
    __shared__ __align__(16) R s_fl[16][4];              // see k_psi_tile
    if (tid < 16) {
        const unsigned f = tid;
        const R nwx = ((f & (NF_MM | NF_MP)) ? (R)1 : (R)0) + ((f & (NF_PM | NF_PP)) ? (R)1 : (R)0);
        const R nwy = ((f & (NF_MM | NF_PM)) ? (R)1 : (R)0) + ((f & (NF_MP | NF_PP)) ? (R)1 : (R)0);
        s_fl[f][0] = (f & (NF_PM | NF_PP)) ? cx : (R)0;
        s_fl[f][1] = (f & (NF_MP | NF_PP)) ? cy : (R)0;
        s_fl[f][2] = idx2 * nwx + idy2 * nwy;
        s_fl[f][3] = f ? (R)1 : (R)0;
    }
    uint32_t phase = 0;

    for (; tile < ntiles; tile += gridDim.x) {
        const int tcur = tile_of(tile);
        const int bx = tcur % ntx, by = tcur / ntx;
        const int xg0 = bx * TX - H, yg0 = g.j0 + by * TYO - K;
        const int x0 = xg0 + c0;                          // global column of the patch's even column
        const int nfd = xg0 - (((xg0 + 1024) / 16) * 16 - 1024);
        if (tid < 32) {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(t_smem_u32(bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        __syncthreads();      // also: everybody is done with the previous tile's planes

        const bool xin[2] = {x0 >= 0 && x0 < g.Nx, x0 + 1 >= 0 && x0 + 1 < g.Nx};
        const bool rhs_is_psi = A.same_rhs && !A.noise;
        if (A.noise) {        // Langevin noise folded into the staged right-hand side (see k_psi_tile)
#pragma unroll 1
            for (int n = 0; n < 2 * PR; n++) {
                const int c = n & 1, r = r0 + (n >> 1), si = r * TXE + c0 + c;
                C qq = A.same_rhs ? ((const C *)(smem + S::st_psi))[si] : ((const C *)(smem + S::st_rhs))[si];
                const unsigned f = xin[c] ? (smem + S::st_nf)[r * NFW + nfd + c0 + c] : 0u;
                if (f) {
                    uint32_t nn = (uint32_t)(x0 + c) + (uint32_t)g.Nx * (uint32_t)(yg0 + r);
                    qq.x += lang * (rand_1<R>(nn, A.rand_t) - (R)0.5);
                    qq.y += lang * (rand_2<R>(nn, A.rand_t) - (R)0.5);
                }
                ((C *)(smem + S::st_rhs))[si] = qq;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        // ---- per-node constants into registers
        C psi[2][PR], q[2][PR], La[2][PR], Lb[2][PR], LbS0[2];
        R di[2][PR];
        LbS0[0].x = LbS0[0].y = LbS0[1].x = LbS0[1].y = 0;
        if (r0 > 0) {         // S-link coefficients of the patch's first row = b-links of tile row r0-1
            const int si = (r0 - 1) * TXE + c0;
            const float2 b01 = *(const float2 *)((const R *)(smem + S::st_b) + si);
            const unsigned f01 = *(const unsigned short *)(smem + S::st_nf + (r0 - 1) * NFW + nfd + c0);
            const unsigned f0 = xin[0] ? (f01 & 0xffu) : 0u, f1 = xin[1] ? (f01 >> 8) : 0u;
            R s0, k0, s1, k1;
            link_sincos2<R, LINKS>(dy * b01.x, dy * b01.y, &s0, &k0, &s1, &k1);
            const R w0 = s_fl[f0][1], w1 = s_fl[f1][1];
            LbS0[0].x = w0 * k0; LbS0[0].y = w0 * s0;
            LbS0[1].x = w1 * k1; LbS0[1].y = w1 * s1;
        }
#pragma unroll
        for (int r = 0; r < PR; r++) {
            const int row = r0 + r, si = row * TXE + c0;
            const float4 p01 = *(const float4 *)((const C *)(smem + S::st_psi) + si);
            const float4 q01 = rhs_is_psi ? p01 : *(const float4 *)((const C *)(smem + S::st_rhs) + si);
            const float2 a01 = *(const float2 *)((const R *)(smem + S::st_a) + si);
            const float2 b01 = *(const float2 *)((const R *)(smem + S::st_b) + si);
            float2 e01;
            if (EPS) e01 = *(const float2 *)((const R *)(smem + S::st_eps) + si);
            else { e01.x = eps0; e01.y = eps0; }
            const unsigned f01 = *(const unsigned short *)(smem + S::st_nf + row * NFW + nfd + c0);
#pragma unroll
            for (int c = 0; c < 2; c++) {
                C p0, qq;
                p0.x = c ? p01.z : p01.x; p0.y = c ? p01.w : p01.y;
                qq.x = c ? q01.z : q01.x; qq.y = c ? q01.w : q01.y;
                const R av = c ? a01.y : a01.x, bv = c ? b01.y : b01.x, e = c ? e01.y : e01.x;
                unsigned f = c ? (f01 >> 8) : (f01 & 0xffu);
                if (!xin[c]) f = 0;
                const float4 w = *(const float4 *)&s_fl[f][0];
                const R wE = w.x, wN = w.y, nw = w.z, act = w.w;
                R sa, ca, sb, cb;
                C la, lb;
                link_sincos2<R, LINKS>(dx * av, dy * bv, &sa, &ca, &sb, &cb);
                la.x = wE * ca; la.y = wE * sa;
                lb.x = wN * cb; lb.y = wN * sb;
                qq.x *= act; qq.y *= act;
                const R D = (R)1.0 + dt * (qq.x * qq.x + qq.y * qq.y - e + nw);
                const R d = f ? rcp_diag(D) : (R)0;       // inactive / out-of-domain nodes stay exactly 0 (see k_psi_tile)
                psi[c][r] = p0; q[c][r] = qq; La[c][r] = la; Lb[c][r] = lb; di[c][r] = d;
            }
            xe[0][(row + 1) * PW + lane] = psi[0][r];
            xo[0][(row + 1) * PW + lane + 1] = psi[1][r];
            slo[(row + 1) * PW + lane + 1] = La[1][r];
        }
        __syncthreads();      // staging fully consumed; level-0 values and the odd columns' link coefficients visible
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);

        unsigned inmask = 0;                  // bit c*4 + r: that node of the patch is an output node of the tile
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int r = 0; r < PR; r++)
                if (c0 + c >= H && c0 + c < TXE - H && x0 + c < g.Nx && r0 + r >= K && r0 + r < EY - K && yg0 + r0 + r < g.j1)
                    inmask |= 1u << (c * 4 + r);

        // the 16 chained FMAs of k_psi_tile, in its order: W, E, S, N
        auto node = [&](const C &qv, const C &lw, const C &pw, const C &le, const C &pe, const C &ls, const C &pS, const C &ln,
                        const C &pN, const R dv) {
            R ax = qv.x, ay = qv.y;
            ax = fma_r(lw.x, pw.x, ax);   ay = fma_r(lw.x, pw.y, ay);
            ax = fma_r(-lw.y, pw.y, ax);  ay = fma_r(lw.y, pw.x, ay);
            ax = fma_r(le.x, pe.x, ax);   ay = fma_r(le.x, pe.y, ay);
            ax = fma_r(le.y, pe.y, ax);   ay = fma_r(-le.y, pe.x, ay);
            ax = fma_r(ls.x, pS.x, ax);   ay = fma_r(ls.x, pS.y, ay);
            ax = fma_r(-ls.y, pS.y, ax);  ay = fma_r(ls.y, pS.x, ay);
            ax = fma_r(ln.x, pN.x, ax);   ay = fma_r(ln.x, pN.y, ay);
            ax = fma_r(ln.y, pN.y, ax);   ay = fma_r(-ln.y, pN.x, ay);
            C o;
            o.x = ax * dv; o.y = ay * dv;
            return o;
        };
        auto sweep = [&](const C(&in)[2][PR], C(&out)[2][PR], const int sb, const int k) {
            const C *se = xe[sb], *so = xo[sb];
            C *de = xe[sb ^ 1], *dd = xo[sb ^ 1];
            const C below0 = se[r0 * PW + lane], below1 = so[r0 * PW + lane + 1];                          // tile row r0-1
            const C above0 = se[(r0 + PR + 1) * PW + lane], above1 = so[(r0 + PR + 1) * PW + lane + 1];   // tile row r0+4
#pragma unroll
            for (int r = 0; r < PR; r++) {
                const int ri = (r0 + r + 1) * PW + lane;
                const C pw0 = so[ri], lw0 = slo[ri];     // column 2l-1 (index l) and its E-link coefficient
                const C pe1 = se[ri + 1];                // column 2l+2 (index l+1)
                out[0][r] = node(q[0][r], lw0, pw0, La[0][r], in[1][r], r > 0 ? Lb[0][r - 1] : LbS0[0],
                                 r > 0 ? in[0][r - 1] : below0, Lb[0][r], r < PR - 1 ? in[0][r + 1] : above0, di[0][r]);
                out[1][r] = node(q[1][r], La[0][r], in[0][r], La[1][r], pe1, r > 0 ? Lb[1][r - 1] : LbS0[1],
                                 r > 0 ? in[1][r - 1] : below1, Lb[1][r], r < PR - 1 ? in[1][r + 1] : above1, di[1][r]);
            }
            const bool more = k < K - 1;
            R rm = 0;
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int r = 0; r < PR; r++)
                    if (inmask & (1u << (c * 4 + r)))
                        rm = fmax(rm, fmax(fabs(out[c][r].x - in[c][r].x), fabs(out[c][r].y - in[c][r].y)));
            if (more) {
#pragma unroll
                for (int r = 0; r < PR; r++) {
                    const int ri = (r0 + r + 1) * PW + lane;
                    de[ri] = out[0][r];
                    dd[ri + 1] = out[1][r];
                }
            }
            const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(rm));
            if (lane == 0 && wm) atomicMax(&sm_rmax[k], wm);
            if (more) __syncthreads();
        };
        {
            C nx[2][PR];
#pragma unroll 1
            for (int k = 0; k + 1 < K; k += 2) {
                sweep(psi, nx, 0, k);
                sweep(nx, psi, 1, k + 1);
            }
            if (K & 1) {
                sweep(psi, nx, 0, K - 1);
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int r = 0; r < PR; r++) psi[c][r] = nx[c][r];
            }
        }
        // ---- write the interior
        {
            C *o = (C *)A.out + ((long long)(yg0 + r0 - g.rb) * g.P + x0);
#pragma unroll
            for (int r = 0; r < PR; r++) {
                if (inmask & (1u << r)) o[(long long)r * g.P] = psi[0][r];
                if (inmask & (1u << (4 + r))) o[(long long)r * g.P + 1] = psi[1][r];
            }
        }
    }
    __syncthreads();
    if (tid < K) {
        const double r = (double)__uint_as_float(sm_rmax[tid]);
        if (r > 0.0) atomicMax(A.slots + tid, (unsigned long long)__double_as_longlong(r));
    }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled t_get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}
static int t_make_map(CUtensorMap *tm, CUtensorMapDataType dt, const void *base, size_t width, size_t rows,
                      size_t pitch_bytes, int box_w, int box_h) {
    PFN_encodeTiled enc = t_get_encode();
    SVL_REQUIRE(enc, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        svl_set_error("cuTensorMapEncodeTiled failed: %d (width %zu rows %zu pitch %zu box %dx%d)", (int)r, width, rows,
                      pitch_bytes, box_w, box_h);
        return 1;
    }
    return 0;
}

// CTAs of the boundary launch of a slab batch: nb boundary tiles, ni interior tiles, `slots` resident CTAs per GPU.
// Boundary CTAs need ceil(nb / t) tile times plus ~3 for their one system-scope fence, the interior ones
// ceil(ni / (slots - t)); the boundary CTAs should be done well before the interior ones (their tiles wait for halo
// flags and the fence time is only roughly known), so a boundary time above 0.85 of the interior time counts as the
// batch time with that margin.  Smallest batch time wins, ties go to the faster boundary.  Pure host arithmetic
// (tests/test_slab_host.py).
extern "C" int svl_slab_split_plan(int nb, int ni, int slots) {
    if (nb <= 0 || slots <= 1) return nb < 1 ? 1 : (nb < slots ? nb : slots);
    long best = -1, best_b = 0;
    int best_t = nb < slots ? nb : slots;
    const int tmax = nb < slots / 2 ? nb : slots / 2;
    for (int t = 1; t <= tmax; t++) {
        const long rb = (nb + t - 1) / t + 3, ri = ((long)ni + (slots - t) - 1) / (slots - t);
        const long r = 20 * rb > 17 * ri ? (rb * 20 + 16) / 17 : ri;
        if (best < 0 || r < best || (r == best && rb < best_b)) { best = r; best_b = rb; best_t = t; }
    }
    return best_t;
}

struct TileIO {
    const void *psi, *rhs, *a, *b, *epsf;
    const uint8_t *nf;
};

// Tensor maps depend only on (base pointer, geometry, box): the three psi buffers rotate, so a
// small cache avoids six driver encode calls per launch.  The cache belongs to the CONTEXT (contexts
// may be driven from different host threads) and forgets a plane when its buffer is freed.
struct MapKey { const void *base; int w, rows, box_w, box_h, dt; size_t pitch; };
struct MapEnt { MapKey k; CUtensorMap tm; };
struct MapCache { MapEnt e[64]; int n, next; };
int svl_tma_map(svl_ctx *c, CUtensorMap *out, CUtensorMapDataType dt, const void *base, size_t width, size_t rows,
                size_t pitch_bytes, int box_w, int box_h) {   // also used by a_tile.cu
    if (!c->tma_cache) c->tma_cache = calloc(1, sizeof(MapCache));
    MapCache *mc = (MapCache *)c->tma_cache;
    SVL_REQUIRE(mc, "out of host memory");
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.w = (int)width; k.rows = (int)rows; k.box_w = box_w; k.box_h = box_h; k.dt = (int)dt; k.pitch = pitch_bytes;
    for (int i = 0; i < mc->n; i++)
        if (!memcmp(&mc->e[i].k, &k, sizeof(k))) { *out = mc->e[i].tm; return 0; }
    CUtensorMap tm;
    SVL_TRY(t_make_map(&tm, dt, base, width, rows, pitch_bytes, box_w, box_h));
    int slot = mc->n < 64 ? mc->n++ : (mc->next++ % 64);
    mc->e[slot].k = k; mc->e[slot].tm = tm;
    *out = tm;
    return 0;
}
// svl_free / svl_destroy: descriptors of a released plane must not outlive it (base == nullptr: all of them)
void svl_tma_forget(svl_ctx *c, const void *base) {
    if (!c || !c->tma_cache) return;
    MapCache *mc = (MapCache *)c->tma_cache;
    if (!base) { free(mc); c->tma_cache = nullptr; return; }
    for (int i = 0; i < mc->n;) {
        if (mc->e[i].k.base == base) mc->e[i] = mc->e[--mc->n];
        else i++;
    }
}

template <typename R, int K, int TXE, int V, int NB, bool EPS, int LINKS>
static int launch_tile_t(svl_ctx *c, TileArgs &A, const TileIO &io) {
    typedef typename V2<R>::type C;
    typedef TileSmem<R, K, TXE, V, NB, EPS> S;
    const Geo &g = c->g;
    constexpr int TX = TXE - 2 * S::H, TYO = S::EY - 2 * K;
    static_assert(TYO > 0 && TX > 0, "tile too small for this K");
    typedef void (*kern_t)(TileArgs, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap);
    kern_t kern = k_psi_tile<R, K, TXE, V, NB, EPS, false, LINKS>;
    auto kern_slab = k_psi_tile<R, K, TXE, V, NB, EPS, true, LINKS>;       // with halo wait + in-kernel push
    if constexpr (sizeof(R) == 4 && TXE == 64 && V == 8 && NB == 4) {
        // fp32: the 2 x 4 patch variant for everything but the boundary tile rows of a slab (option psi_patch)
        static bool patch_ready = false;
        if (!patch_ready) {
            SVL_CHECK(cudaFuncSetAttribute(k_psi_patch<K, EPS, LINKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
            patch_ready = true;
        }
        if (c->opt_psi_patch) kern = k_psi_patch<K, EPS, LINKS>;
    }
    static int slots = 0;
    if (!slots) {
        int occ = 1, nsm = 148;
        SVL_CHECK(cudaFuncSetAttribute(k_psi_tile<R, K, TXE, V, NB, EPS, false, LINKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        SVL_CHECK(cudaFuncSetAttribute(kern_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        SVL_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_slab, TXE * NB, S::total));
        SVL_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
        slots = (occ < 1 ? 1 : occ) * nsm;
    }
    const bool dbl = sizeof(R) == 8;
    CUtensorMapDataType rt = dbl ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMapDataType ct = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    const int cmul = dbl ? 2 : 1;
    static_assert(TXE * 2 <= 256, "complex double box exceeds 256 elements");
    size_t pr = (size_t)g.P * sizeof(R), pc = (size_t)g.P * sizeof(C);
    CUtensorMap tm[6];
    SVL_TRY(svl_tma_map(c, &tm[0], ct, io.psi, (size_t)g.Nx * cmul, g.rows, pc, TXE * cmul, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[1], ct, io.rhs, (size_t)g.Nx * cmul, g.rows, pc, TXE * cmul, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[2], rt, io.a, g.Nx, g.rows, pr, TXE, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[3], rt, io.b, g.Nx, g.rows, pr, TXE, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[4], rt, EPS ? io.epsf : io.a, g.Nx, g.rows, pr, TXE, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[5], CU_TENSOR_MAP_DATA_TYPE_UINT8, io.nf, g.Nx, g.rows, (size_t)g.P, S::NFW, S::EY));
    const int ntx_ = (g.Nx + TX - 1) / TX, nty_ = (g.j1 - g.j0 + TYO - 1) / TYO, rows_ = g.j1 - g.j0;
    A.push_expect[0] = A.push_expect[1] = 0;
    int nlo = 0, nhi = 0;                                  // tile rows that feed the lo / hi push
    for (int by = 0; by < nty_; by++) {
        if (by * TYO < A.push.depth) { A.push_expect[0] += ntx_; if (A.push.peer[0][0]) nlo++; }
        if (((by + 1) * TYO < rows_ ? (by + 1) * TYO : rows_) > rows_ - A.push.depth) { A.push_expect[1] += ntx_; if (A.push.peer[1][0]) nhi++; }
    }
    A.row0 = 0; A.nrow = nty_; A.row1 = 0; A.nrow1 = 0; A.permute = 0;
    if (!A.wait_flags) {
        int ntiles = ntx_ * nty_;
        if (c->opt_pdl && !A.gate) {
            // Consecutive batches of a solve as programmatic dependent launches: a batch releases its successor at its
            // first instruction, the successor's CTAs become resident as this batch's CTAs retire, set up their shared
            // memory and then wait (griddepcontrol.wait, thread 0, before the first TMA load) for this batch to end: the
            // launch latency, the prologue and part of the tail of every batch are hidden.
            A.pdl_trigger = 1; A.pdl_wait = 1;
            cudaLaunchConfig_t lc;
            memset(&lc, 0, sizeof(lc));
            lc.gridDim = dim3(ntiles < slots ? ntiles : slots); lc.blockDim = dim3(TXE * NB); lc.dynamicSmemBytes = S::total;
            lc.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = at; lc.numAttrs = 1;
            SVL_CHECK(cudaLaunchKernelEx(&lc, kern, A, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]));
        } else {
            kern<<<ntiles < slots ? ntiles : slots, TXE * NB, S::total, c->stream>>>(A, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]);
        }
    } else if (c->opt_slab_split && nty_ - nlo - nhi > 0 && nlo + nhi > 0) {
        // Slabs, two launches per batch ON ONE STREAM: first the boundary tile rows (halo wait + in-kernel push; at
        // most a few dozen CTAs, resident at once), then the interior rows with the plain kernel as a PROGRAMMATIC
        // DEPENDENT launch: the boundary kernel releases it at its first instruction (griddepcontrol.launch_dependents),
        // so both run side by side without any event traffic between streams (round 1 used a fork/join of two streams:
        // ~11 us per batch, 12 % of a 85 us launch).  The two kernels read `cur` and write disjoint rows of `out`; the
        // interior kernel consumes nothing of the boundary kernel, so it never waits on it.  The next batch is an
        // ordinary launch and therefore starts after both have completed.
        TileArgs B = A;
        B.row0 = 0; B.nrow = nlo; B.row1 = nty_ - nhi; B.nrow1 = nhi;
        B.pdl_trigger = 1;
        const int nb = ntx_ * (nlo + nhi), ni = ntx_ * (nty_ - nlo - nhi);
        // The boundary tiles go to a FEW CTAs (each takes several tiles and fences once), the interior kernel gets the
        // remaining slots, all resident from the start: a boundary CTA holds its slot through the system-scope fence,
        // and an interior CTA that had to wait for that slot would finish its static share of tiles late.  The batch
        // lasts max(rounds of the interior CTAs, rounds of the boundary CTAs + ~3 tile times of fence) tile times; both
        // are step functions of the split (2048^2 per GPU with two neighbours: 13 boundary CTAs keep the interior at
        // 11 rounds like a single GPU, 15 would push it to 12), so the split is chosen by evaluating that maximum.
        int gb = nb < slots ? nb : slots;
        if (c->opt_slab_bnd > 0) gb = c->opt_slab_bnd < gb ? c->opt_slab_bnd : gb;
        else if (c->opt_slab_bnd == 0) gb = svl_slab_split_plan(nb, ni, slots);
        B.defer_publish = 1;
        const bool chain = c->opt_pdl >= 2 && !A.gate;
        if (chain) {
            // Batches chained across the pair of launches: this boundary launch is a programmatic dependent of the previous
            // batch's INTERIOR launch (which released it at its first instruction and does not complete before its own
            // boundary launch has, pdl_wait_end); its CTAs set up, wait for that batch to end, and only then release this
            // batch's interior launch.
            B.pdl_wait = 1; B.pdl_trigger = 2;
            cudaLaunchConfig_t lb;
            memset(&lb, 0, sizeof(lb));
            lb.gridDim = dim3(gb); lb.blockDim = dim3(TXE * NB); lb.dynamicSmemBytes = S::total; lb.stream = c->stream;
            cudaLaunchAttribute ab[1];
            ab[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            ab[0].val.programmaticStreamSerializationAllowed = 1;
            lb.attrs = ab; lb.numAttrs = 1;
            SVL_CHECK(cudaLaunchKernelEx(&lb, kern_slab, B, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]));
        } else {
            kern_slab<<<gb, TXE * NB, S::total, c->stream>>>(B, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]);
        }
        SVL_CHECK(cudaGetLastError());
        TileArgs I = A;
        I.wait_flags = nullptr; memset(&I.push, 0, sizeof(I.push)); I.trace = nullptr;
        if (chain) { I.pdl_trigger = 1; I.pdl_wait_end = 1; }
        I.row0 = nlo; I.nrow = nty_ - nlo - nhi;
        int gi = c->opt_slab_bnd < 0 ? slots : slots - gb;
        if (gi < 1) gi = 1;
        if (gi > ni) gi = ni;
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3(gi); lc.blockDim = dim3(TXE * NB); lc.dynamicSmemBytes = S::total; lc.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        SVL_CHECK(cudaLaunchKernelEx(&lc, kern, I, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]));
        c->stat_launches += 1;
    } else {
        A.permute = 1;                                   // boundary rows first, then the interior
        A.row0 = 0; A.nrow = nlo; A.row1 = nty_ - nhi; A.nrow1 = nhi;
        if (nty_ - nlo - nhi < 0) { A.permute = 0; A.row0 = 0; A.nrow = nty_; A.row1 = 0; A.nrow1 = 0; }
        int ntiles = ntx_ * nty_;
        kern_slab<<<ntiles < slots ? ntiles : slots, TXE * NB, S::total, c->stream>>>(A, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]);
    }
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

template <typename R, int TXE, int V, int NB, bool EPS, int LINKS>
static int launch_tile_k(svl_ctx *c, int K, TileArgs &A, const TileIO &io) {
    switch (K) {
        case 1: return launch_tile_t<R, 1, TXE, V, NB, EPS, LINKS>(c, A, io);
        case 2: return launch_tile_t<R, 2, TXE, V, NB, EPS, LINKS>(c, A, io);
        case 3: return launch_tile_t<R, 3, TXE, V, NB, EPS, LINKS>(c, A, io);
        case 4: return launch_tile_t<R, 4, TXE, V, NB, EPS, LINKS>(c, A, io);
        case 6: return launch_tile_t<R, 6, TXE, V, NB, EPS, LINKS>(c, A, io);
        case 8: return launch_tile_t<R, 8, TXE, V, NB, EPS, LINKS>(c, A, io);
    }
    svl_set_error("psi_tile: K=%d not instantiated (1,2,3,4,6,8)", K);
    return 2;
}

int svl_launch_psi_tile(svl_ctx *c, int K, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                        const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t rand_t,
                        unsigned long long *resid_slots) {
    TileArgs A;
    memset(&A, 0, sizeof(A));
    A.g = c->g;
    A.dt = dt; A.eps = eps; A.lang_c = lang_c; A.rand_t = rand_t;
    A.noise = lang_c > 1.0e-32 ? 1 : 0;
    A.same_rhs = rhs->p[0] == psi->p[0];
    A.out = out->p[0]; A.slots = resid_slots;
    A.sg = svl_spin_guard(c);
    A.gate = c->spec_gate;
    if (c->slab_on && !c->opt_slab_nocomm) {
        A.wait_flags = c->flags; A.wait_epoch = svl_slab_epoch(c); A.has_lo = c->has_lo; A.has_hi = c->has_hi;
        svl_slab_mark_waited(c);
        SVL_TRY(svl_slab_push_fused(c, out, &A.push));          // after wait_epoch: this launch's own push
        if (c->trace && c->trace_n < c->trace_cap) {
            A.trace = c->trace + 4 * (size_t)c->trace_n;
            c->trace_n += 1;
        }
    }
    TileIO io = {psi->p[0], rhs->p[0], ab->p[0], ab->p[1], epsf ? epsf->p[0] : nullptr, c->nf};
    if (c->rsize == 4) {
        // fp32 thread shape (option psi_shape): 0 = 512 threads x 4 rows each at 64 registers, two CTAs = 32 warps per
        // SM; 1 = 256 threads x 8 rows each at 128 registers, 16 warps per SM (fewer instructions per node, half the
        // warps to hide latencies and barriers with)
        const int lk = c->opt_psi_links ? 1 : 0, sh = c->opt_psi_shape ? 1 : 0, ep = epsf ? 1 : 0;
        // (64 x 64 tiles -- 77 % instead of 66 % of the computed nodes are output -- on 512 threads, one CTA per SM, were
        // measured as well: 9 % fewer instructions, 0.370 vs 0.365 ms/step; one CTA per SM leaves nobody to issue while
        // its warps sit at the sweep barrier, and 1369 tiles on 148 CTAs quantise worse than 3182 on 296)
        switch (sh * 4 + lk * 2 + ep) {
            case 0: return launch_tile_k<float, 64, 4, 8, false, 0>(c, K, A, io);
            case 1: return launch_tile_k<float, 64, 4, 8, true, 0>(c, K, A, io);
            case 2: return launch_tile_k<float, 64, 4, 8, false, 1>(c, K, A, io);
            case 3: return launch_tile_k<float, 64, 4, 8, true, 1>(c, K, A, io);
            case 4: return launch_tile_k<float, 64, 8, 4, false, 0>(c, K, A, io);
            case 5: return launch_tile_k<float, 64, 8, 4, true, 0>(c, K, A, io);
            case 6: return launch_tile_k<float, 64, 8, 4, false, 1>(c, K, A, io);
            default: return launch_tile_k<float, 64, 8, 4, true, 1>(c, K, A, io);
        }
    }
    // (fp64 with 256 threads x 8 rows at 244 registers, 8 warps per SM, was measured: cfg3 18.9 -> 19.9 ms/step)
    if (epsf) return launch_tile_k<double, 64, 4, 8, true, 0>(c, K, A, io);
    return launch_tile_k<double, 64, 4, 8, false, 0>(c, K, A, io);
}
