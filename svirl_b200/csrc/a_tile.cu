// Jacobi sweeps of the vector-potential equation, two sweeps per launch on an overlapped 2-D tile
// (same arithmetic as k_a_sweep in td.cu / svirl/cuda/td.h:311-463).
//
// Why pairs: the reference evaluates the supercurrent link phase of sweep s from Jacobi iterate
// s - (s mod 2) (quirk Q1, svirl/solvers/td.py:266,286,303).  For an even sweep s that iterate IS
// the sweep's input, and the following odd sweep uses the same one -- so the supercurrent
// Im(conj(psi0) U(d A) psi1)/d, the Langevin noise, the boundary field term and the right-hand
// side are one constant per edge for the two sweeps of a pair.  A launch therefore
//   * stages psi, a, b of the extended tile (interior + halo of K rows / H columns) into shared
//     memory with three 2-D TMA boxes (zero fill outside the plane = domain boundary),
//   * evaluates  c = rhs + dt*rho*(js + rh)  once per edge (one sincos per edge and PAIR instead
//     of one per edge and sweep; never stored in global memory),
//   * runs K (1 or 2) Jacobi sweeps out of shared memory / registers (each thread owns V rows of
//     one column; only the W/E neighbour columns and the strip ends are read from shared memory),
//   * writes the interior once and reduces the max-norm update of each sweep (one atomicMax per
//     CTA and sweep on the bit pattern: exact, order independent).
// HBM traffic per launch: psi 2R + a,b 2R (x halo overhead) + rhs 2R + flags 1 + out 2R = 8R+1 per
// node for two sweeps, against 2 x (9R+1) for two single sweeps.
#include "common.cuh"
#include <cuda.h>
#include <type_traits>

int svl_tma_map(svl_ctx *c, CUtensorMap *out, CUtensorMapDataType dt, const void *base, size_t width, size_t rows,
                size_t pitch_bytes, int box_w, int box_h);   // psi_tile.cu

struct ATileArgs {
    Geo g;
    double dt, kappa2, rho, H, lang_c;
    uint32_t rand_t;
    int noise;
    const void *rhs_a, *rhs_b;
    const uint8_t *nf;
    void *out_a, *out_b;
    unsigned long long *slots;
    const unsigned long long *wait_flags;
    unsigned long long wait_epoch;
    SpinGuard sg;                  // bound of the spin waits on the neighbours' flags
    int has_lo, has_hi;
    SlabPush push;                 // slabs: in-kernel push of the output's boundary rows
    int push_expect[2];
};

__device__ __forceinline__ uint32_t a_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(a_smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(a_smem_u32(bar))
        : "memory");
}

// Shared memory: [guard][psi C x NN][a0][b0][guard][a1][b1][guard][barrier].  Tile-edge nodes read one
// row / column beyond their array; those reads stay inside this block (guards) and only feed
// tile-edge results, which are never used.  The guard between b0 and a1 keeps such a read of the first sweep
// away from the words the same sweep writes (a benign but real read/write hazard that compute-sanitizer's
// racecheck reported in round 2).
template <typename R, int TXE, int V, int NB>
struct ASmem {
    typedef typename V2<R>::type C;
    static constexpr int EY = V * NB;
    static constexpr int NN = EY * TXE;
    static constexpr size_t guard = 1024;                       // >= (TXE + 1) * sizeof(R), multiple of 128
    static constexpr size_t off_psi = guard;
    static constexpr size_t off_a0 = off_psi + sizeof(C) * NN;
    static constexpr size_t off_b0 = off_a0 + sizeof(R) * NN;
    static constexpr size_t off_a1 = off_b0 + sizeof(R) * NN + guard;
    static constexpr size_t off_b1 = off_a1 + sizeof(R) * NN;
    static constexpr size_t off_bar = off_b1 + sizeof(R) * NN + guard;
    static constexpr size_t total = off_bar + 16;
    static constexpr uint32_t tx_bytes = (uint32_t)((sizeof(C) + 2 * sizeof(R)) * NN);
};

template <typename R, int K, int TXE, int V, int NB, int MODE, bool SLAB>
__global__ void __launch_bounds__(TXE *NB, (TXE * NB > 256 ? 1 : 2))
k_a_tile(const __grid_constant__ ATileArgs A, const __grid_constant__ CUtensorMap tm_psi,
         const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b) {
    typedef typename V2<R>::type C;
    typedef ASmem<R, TXE, V, NB> S;
    constexpr int EY = S::EY;
    constexpr int H = sizeof(R) == 8 ? 2 : 4;             // halo columns: boxes start on 16-byte boundaries
    constexpr int TX = TXE - 2 * H, TYO = EY - 2 * K;
    extern __shared__ __align__(128) unsigned char smem[];
    C *spsi = (C *)(smem + S::off_psi);                    // psi, later the per-edge constants (ca, cb)
    R *sa0 = (R *)(smem + S::off_a0), *sb0 = (R *)(smem + S::off_b0);
    R *sa1 = (R *)(smem + S::off_a1), *sb1 = (R *)(smem + S::off_b1);
    uint64_t *bar = (uint64_t *)(smem + S::off_bar);
    __shared__ unsigned long long sm_rmax[2];

    const Geo &g = A.g;
    const int tid = threadIdx.x;
    const int col = tid % TXE, band = tid / TXE;
    const int r0 = band * V;
    const int ntx = (g.Nx + TX - 1) / TX, nty = (g.j1 - g.j0 + TYO - 1) / TYO;
    // slabs: the first / last tile row (the only tiles that read a neighbour's halo rows) get the
    // lowest block indices: they run first (the neighbour's previous push landed a launch ago, so the
    // flag wait is free) and this launch's push leaves early
    int tile = blockIdx.x;
    if (SLAB && nty > 2) tile = tile < ntx ? tile : (tile < 2 * ntx ? (nty - 1) * ntx + (tile - ntx) : tile - ntx);
    const int bx = tile % ntx, by = tile / ntx;
    const int xg0 = bx * TX - H, yg0 = g.j0 + by * TYO - K;
    const int x = xg0 + col;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (SLAB) {                                        // slabs: the neighbours' halo rows of the input have arrived
            for (int sdir = 0; sdir < 2; sdir++) {
                if (!(sdir == 0 ? A.has_lo : A.has_hi)) continue;
                const int rows_own = g.j1 - g.j0, top = (by + 1) * TYO < rows_own ? (by + 1) * TYO : rows_own;
                if (sdir == 0 ? (by * TYO - K >= 0) : (top + K <= rows_own)) continue;   // box stays inside the slab
                svl_spin_ge(A.wait_flags + sdir, A.wait_epoch, A.sg);
            }
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a_smem_u32(bar)), "r"(S::tx_bytes) : "memory");
        const int cmul = sizeof(C) == 16 ? 2 : 1;
        const int prow = yg0 - g.rb;
        a_tma_load_2d(spsi, &tm_psi, xg0 * cmul, prow, bar);
        a_tma_load_2d(sa0, &tm_a, xg0, prow, bar);
        a_tma_load_2d(sb0, &tm_b, xg0, prow, bar);
        sm_rmax[0] = 0ull; sm_rmax[1] = 0ull;
    }

    // ---- own-edge data straight from global memory (coalesced along x), overlapping the TMA
    const R dx = (R)g.dx, dy = (R)g.dy, idx = (R)g.idx, idy = (R)g.idy, idx2 = (R)g.idx2, idy2 = (R)g.idy2,
            idxy = (R)g.idxy;
    const R dt_rho = (R)A.dt * (R)A.rho, dtrk = dt_rho * (R)A.kappa2;
    const R inv_da = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idy2), inv_db = (R)1.0 / ((R)1.0 + (R)2.0 * dtrk * idx2);
    const R kH2 = (R)2.0 * (R)A.kappa2 * (R)A.H;
    const bool xin = x >= 0 && x < g.Nx;
    // c starts as the right-hand side q (own edges, straight from global memory: coalesced along x,
    // overlapping the TMA) and becomes  c = q + dt*rho*(js + rh)  below
    R ca[V], cb[V];
    unsigned fl[V];
    auto load_rhs = [&]() {
#pragma unroll
        for (int v = 0; v < V; v++) {
            const int pr = yg0 + r0 + v - g.rb;
            ca[v] = 0; cb[v] = 0; fl[v] = 0;
            if (xin && pr >= 0 && pr < g.rows) {
                const size_t n = (size_t)pr * g.P + x;
                ca[v] = ((const R *)A.rhs_a)[n];
                cb[v] = ((const R *)A.rhs_b)[n];
                fl[v] = A.nf[n];
            }
        }
    };
    load_rhs();
    // ---- wait for the boxes: one warp polls, the barrier releases the rest
    if (tid < 32) {
        uint32_t ok = 0;       // no iteration bound: thread 0 may still be waiting for a neighbour's halo flag
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(a_smem_u32(bar)) : "memory");
    }
    __syncthreads();

    // ---- per-edge constants  c = q + dt*rho*(js + rh)   (td.h:366-405, 412-451)
    // boundary doubling of the b-edge depends on the column only
    const R rh_b = x == 0 ? -kH2 * idx : (x == g.Nx - 1 ? kH2 * idx : (R)0);
    const R wb = dtrk * ((x == 0 || x == g.Nx - 1) ? (R)2 : (R)1);
    // The fast pass evaluates all 2V link variables with the branch-free sincos core (the independent
    // polynomials interleave) and notes whether any phase was outside its range; only then the
    // constants are redone through libdevice.
    bool bad = false;
    auto constants = [&](auto fastpath) {
#pragma unroll
        for (int v = 0; v < V; v++) {
            const int r = r0 + v, si = r * TXE + col, y = yg0 + r;
            const C p0 = spsi[si], pE = spsi[si + 1], pN = spsi[si + TXE];
            const R pha = dx * sa0[si], phb = dy * sb0[si];
            if (A.noise) {      // Langevin term: added to the right-hand side on sweep 0 (td.h:370-373, 416-419)
                const R lang = (R)A.lang_c;
                const uint32_t na = (uint32_t)x + (uint32_t)(g.Nx - 1) * (uint32_t)y;
                const uint32_t nb = (uint32_t)((size_t)(g.Nx - 1) * g.Ny) + (uint32_t)x + (uint32_t)g.Nx * (uint32_t)y;
                ca[v] += lang * (rand_1<R>(na, A.rand_t) - (R)0.5);
                cb[v] += lang * (rand_2<R>(nb, A.rand_t) - (R)0.5);
            }
            R sa, cA, sb, cB;
            if (decltype(fastpath)::value == 1) {
                bad = bad || !sincos_fast_ok(pha) || !sincos_fast_ok(phb);
                sincos_fast(pha, &sa, &cA); sincos_fast(phb, &sb, &cB);
            } else if (decltype(fastpath)::value == 2) {       // range test per link (no interleaving)
                sincos_r<R>(pha, &sa, &cA); sincos_r<R>(phb, &sb, &cB);
            } else {
                sincos_any(pha, &sa, &cA); sincos_any(phb, &sb, &cB);
            }
            // Im(conj(p0) U(ph) p1) = (p0 x p1) cos - (p0 . p1) sin   (svirl/cuda/common.h:65-73)
            R jla = idx * ((p0.x * pE.y - p0.y * pE.x) * cA - (p0.x * pE.x + p0.y * pE.y) * sa);
            R jlb = idy * ((p0.x * pN.y - p0.y * pN.x) * cB - (p0.x * pN.x + p0.y * pN.y) * sb);
            if (!(fl[v] & (NF_PM | NF_PP))) jla = 0;
            if (!(fl[v] & (NF_MP | NF_PP))) jlb = 0;
            const R rh_a = y == 0 ? kH2 * idy : (y == g.Ny - 1 ? -kH2 * idy : (R)0);
            ca[v] = ca[v] + dt_rho * (jla + rh_a);
            cb[v] = cb[v] + dt_rho * (jlb + rh_b);
        }
    };
    if (MODE == 1) {
        constants(std::true_type());
        if (bad) { load_rhs(); constants(std::false_type()); }
    } else {
        constants(std::integral_constant<int, 2>());
    }
    R Av[V], Bv[V];
#pragma unroll
    for (int v = 0; v < V; v++) { Av[v] = sa0[(r0 + v) * TXE + col]; Bv[v] = sb0[(r0 + v) * TXE + col]; }

    const bool cin = col >= H && col < TXE - H && x < g.Nx;
    unsigned inmask = 0;                 // bit v: node v of this thread is an output node of the tile
#pragma unroll
    for (int v = 0; v < V; v++)
        if (cin && r0 + v >= K && r0 + v < EY - K && yg0 + r0 + v < g.j1) inmask |= 1u << v;

    const R *srcA = sa0, *srcB = sb0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        R rm = 0;
        const R belowA = srcA[(r0 - 1) * TXE + col], belowB = srcB[(r0 - 1) * TXE + col];
        const R aboveA = srcA[(r0 + V) * TXE + col];
        R bE_prev = srcB[(r0 - 1) * TXE + col + 1];          // B(r-1, c+1)
        R aW = srcA[r0 * TXE + col - 1];                     // A(r, c-1)
        R na[V], nb[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
            const int r = r0 + v, si = r * TXE + col, y = yg0 + r;
            const R Am = v > 0 ? Av[v - 1] : belowA, Ap = v < V - 1 ? Av[v + 1] : aboveA;
            const R Bm = v > 0 ? Bv[v - 1] : belowB;
            const R bE = srcB[si + 1];                       // B(r, c+1)
            const R bW = srcB[si - 1];                       // B(r, c-1)
            const R aWp = srcA[si + TXE - 1];                // A(r+1, c-1)
            // a-edge (td.h:375-405)
            const bool va = x >= 0 && x < g.Nx - 1 && y >= 0 && y < g.Ny;
            const R wa = dtrk * ((y == 0 || y == g.Ny - 1) ? (R)2 : (R)1);
            R lo = idy2 * Am - idxy * Bm + idxy * bE_prev;
            R hi = idy2 * Ap + idxy * Bv[v] - idxy * bE;
            R xa = (ca[v] + wa * (lo + hi)) * inv_da;
            na[v] = va ? xa : (R)0;
            // b-edge (td.h:421-451)
            const bool vb = xin && y >= 0 && y < g.Ny - 1;
            lo = idx2 * bW - idxy * aW + idxy * aWp;
            hi = idx2 * bE + idxy * Av[v] - idxy * Ap;
            R xb = (cb[v] + wb * (lo + hi)) * inv_db;
            nb[v] = vb ? xb : (R)0;
            bE_prev = bE;
            aW = aWp;
        }
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (inmask & (1u << v)) rm = fmax(rm, fmax(fabs(na[v] - Av[v]), fabs(nb[v] - Bv[v])));
            Av[v] = na[v]; Bv[v] = nb[v];
            if (k < K - 1) { sa1[(r0 + v) * TXE + col] = na[v]; sb1[(r0 + v) * TXE + col] = nb[v]; }
        }
        for (int o = 16; o > 0; o >>= 1) rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o));
        if ((tid & 31) == 0 && rm > (R)0) atomicMax(&sm_rmax[k], (unsigned long long)__double_as_longlong((double)rm));
        if (k < K - 1) __syncthreads();
        srcA = sa1; srcB = sb1;
    }
    // ---- write the interior (slabs: the first / last `depth` rows also go to the neighbours' halos)
    const int rows_own = g.j1 - g.j0;
    const bool plo = SLAB && A.push.peer[0][0] && by * TYO < A.push.depth;
    const bool phi = SLAB && A.push.peer[1][0] && ((by + 1) * TYO < rows_own ? (by + 1) * TYO : rows_own) > rows_own - A.push.depth;
#pragma unroll
    for (int v = 0; v < V; v++) {
        if (inmask & (1u << v)) {
            const int y = yg0 + r0 + v;
            const size_t n = g.at(x, y);
            ((R *)A.out_a)[n] = Av[v];
            ((R *)A.out_b)[n] = Bv[v];
            if (plo && y < g.j0 + A.push.depth) {
                const size_t m = (size_t)(y - A.push.peer_rb[0]) * g.P + x;
                ((R *)A.push.peer[0][0])[m] = Av[v]; ((R *)A.push.peer[0][1])[m] = Bv[v];
            }
            if (phi && y >= g.j1 - A.push.depth) {
                const size_t m = (size_t)(y - A.push.peer_rb[1]) * g.P + x;
                ((R *)A.push.peer[1][0])[m] = Av[v]; ((R *)A.push.peer[1][1])[m] = Bv[v];
            }
        }
    }
    __syncthreads();                             // all peer stores of the CTA issued ...
    if ((plo || phi) && tid == 0) {
        __threadfence_system();                  // ... and made visible by ONE fence (grid-sync pattern)
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
    if (tid < K) {
        unsigned long long b = sm_rmax[tid];
        if (b) atomicMax(A.slots + tid, b);
    }
}

template <typename R, int K, int TXE, int V, int NB, int MODE>
static int launch_a_tile_t(svl_ctx *c, ATileArgs &A, const void *psi, const void *a, const void *b) {
    typedef typename V2<R>::type C;
    typedef ASmem<R, TXE, V, NB> S;
    const Geo &g = c->g;
    constexpr int H = sizeof(R) == 8 ? 2 : 4;
    constexpr int TX = TXE - 2 * H, TYO = S::EY - 2 * K;
    static_assert(S::guard >= (TXE + 1) * sizeof(R), "guard too small");
    auto kern = k_a_tile<R, K, TXE, V, NB, MODE, false>;
    auto kern_slab = k_a_tile<R, K, TXE, V, NB, MODE, true>;       // with halo wait + in-kernel push
    static bool configured = false;
    if (!configured) {
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        SVL_CHECK(cudaFuncSetAttribute(kern_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        configured = true;
    }
    const bool dbl = sizeof(R) == 8;
    CUtensorMapDataType rt = dbl ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const int cmul = dbl ? 2 : 1;
    size_t pr = (size_t)g.P * sizeof(R), pc = (size_t)g.P * sizeof(C);
    CUtensorMap tm[3];
    SVL_TRY(svl_tma_map(c, &tm[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, psi, (size_t)g.Nx * cmul, g.rows, pc, TXE * cmul, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[1], rt, a, g.Nx, g.rows, pr, TXE, S::EY));
    SVL_TRY(svl_tma_map(c, &tm[2], rt, b, g.Nx, g.rows, pr, TXE, S::EY));
    const int ntx_ = (g.Nx + TX - 1) / TX, nty_ = (g.j1 - g.j0 + TYO - 1) / TYO, rows_ = g.j1 - g.j0;
    A.push_expect[0] = A.push_expect[1] = 0;
    for (int by = 0; by < nty_; by++) {
        if (by * TYO < A.push.depth) A.push_expect[0] += ntx_;
        if (((by + 1) * TYO < rows_ ? (by + 1) * TYO : rows_) > rows_ - A.push.depth) A.push_expect[1] += ntx_;
    }
    int ntiles = ntx_ * nty_;
    if (A.wait_flags) kern_slab<<<ntiles, TXE * NB, S::total, c->stream>>>(A, tm[0], tm[1], tm[2]);
    else kern<<<ntiles, TXE * NB, S::total, c->stream>>>(A, tm[0], tm[1], tm[2]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

// K = 1 or 2 sweeps starting at an EVEN sweep index (link phase = input iterate).
int svl_launch_a_tile(svl_ctx *c, int K, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                      const svl_buf *rhs, const svl_buf *ab, svl_buf *out, double lang_c, uint32_t rand_t,
                      unsigned long long *resid_slots) {
    SVL_REQUIRE(K == 1 || K == 2, "a_tile: K must be 1 or 2");
    SVL_REQUIRE(out->p[0] != ab->p[0] && out->p[0] != rhs->p[0], "a_tile: output must not alias the input or the right-hand side");
    ATileArgs A;
    memset(&A, 0, sizeof(A));
    A.g = c->g;
    A.dt = dt; A.kappa2 = kappa2; A.rho = rho; A.H = H; A.lang_c = lang_c; A.rand_t = rand_t;
    A.noise = lang_c > 1.0e-32 ? 1 : 0;
    A.rhs_a = rhs->p[0]; A.rhs_b = rhs->p[1];
    A.nf = c->nf;
    A.out_a = out->p[0]; A.out_b = out->p[1];
    A.slots = resid_slots;
    if (c->slab_on && !c->opt_slab_nocomm) {
        A.wait_flags = c->flags; A.wait_epoch = svl_slab_epoch(c); A.has_lo = c->has_lo; A.has_hi = c->has_hi;
        A.sg = svl_spin_guard(c);
        svl_slab_mark_waited(c);
        SVL_TRY(svl_slab_push_fused(c, out, &A.push));          // after wait_epoch: this launch's own push
    }
    const void *P_ = psi->p[0], *a_ = ab->p[0], *b_ = ab->p[1];
    // a_kernel option: 1 = 64x32 tile, 8 rows per thread, 2 CTAs/SM (default); 2 = same with the
    // interleaved sincos pass; 3 / 4 = 4 rows per thread, 512 threads (per-link / interleaved)
#define A_TILE_CASE(R_, K_) \
    switch (c->opt_a_kernel) { \
        case 2: return launch_a_tile_t<R_, K_, 64, 8, 4, 1>(c, A, P_, a_, b_); \
        case 3: return launch_a_tile_t<R_, K_, 64, 4, 8, 0>(c, A, P_, a_, b_); \
        case 4: return launch_a_tile_t<R_, K_, 64, 4, 8, 1>(c, A, P_, a_, b_); \
        default: return launch_a_tile_t<R_, K_, 64, 8, 4, 0>(c, A, P_, a_, b_); \
    }
    if (c->rsize == 4) {
        if (K == 2) A_TILE_CASE(float, 2)
        A_TILE_CASE(float, 1)
    }
    if (K == 2) A_TILE_CASE(double, 2)
    A_TILE_CASE(double, 1)
#undef A_TILE_CASE
}
