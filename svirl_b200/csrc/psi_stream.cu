// Temporally blocked, streaming Jacobi sweeps of the psi equation: K sweeps per launch.
//
// Same arithmetic as k_psi_sweep (td.cu; reference svirl/cuda/td.h:5-133), restructured for
// the B200 memory system:
//   * a CTA owns an x-strip of TX = EX - 2H output columns (one thread per column of the strip
//     extended by H = roundup(K, 4) halo columns on each side, so that every TMA box starts on a
//     16-byte boundary) and a y-segment of TY output rows, and marches through the rows once
//     ("2.5-D streaming");
//   * the K sweeps form a software pipeline skewed by two rows per level: when input row r
//     arrives, sweep k (k = 1..K) produces row r - 2k from rows r-2k-1 .. r-2k+1 of sweep k-1.
//     Every level keeps a 4-row ring in shared memory, so each input element is read from HBM
//     once per launch and each output element written once: HBM traffic per node and launch is
//     the 8R+1(+R) bytes of ONE sweep, for K sweeps;
//   * the link variables exp(-i d A) are evaluated once per link and launch (2 sincos per node
//     instead of 4 per node and sweep), pre-multiplied by weight*dt/d^2 and kept in a shared
//     memory ring of 2K+2 rows; they never go to global memory;
//   * input rows (psi, rhs, a, b, flags, eps) are staged into shared memory by 2-D TMA boxes
//     (cp.async.bulk.tensor, mbarrier completion, zero fill outside the plane) issued by one
//     thread a few rows ahead of the consumers; a plain-load staging path (TMA = false) exists
//     for bring-up and as a cross-check;
//   * the max-norm update of every one of the K sweeps is reduced per CTA and merged with one
//     atomicMax on the bit pattern per sweep (exact and order independent), so the host can
//     find the reference's exact stop sweep.
#include "common.cuh"
#include <cuda.h>

#define PS_RD 4   // ring rows per level

template <typename R> struct R2T;
template <> struct R2T<float>  { typedef float2 type; };
template <> struct R2T<double> { typedef double2 type; };

struct StreamArgs {
    Geo g;
    double dt, eps, lang_c;
    uint32_t rand_t;
    int TY;            // output rows per segment
    int has_eps;
    int noise;
    int same_rhs;      // rhs and psi are the same buffer (first launch of a solve): load it once
    const void *psi, *rhs, *a, *b, *epsf;
    const uint8_t *nf;
    void *out;
    unsigned long long *slots;
};

// ------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------- shared memory plan
template <typename R, int K, int EX, int RB, int NS, bool EPS>
struct StreamSmem {
    typedef typename V2<R>::type C;
    static constexpr int CD = 2 * K + 2;          // constants ring rows
    static constexpr int W = EX + 2;              // ring row width (one pad column each side)
    static constexpr int H = ((K + 3) / 4) * 4;   // x halo: >= K and a multiple of 4 columns (16-byte box starts)
    static constexpr int NFW = EX + 16;           // flag rows are fetched from a 16-aligned start, 16 bytes wider
    // staging: NS stages x RB rows x EX columns of every input field
    static constexpr size_t st_psi = 0;
    static constexpr size_t st_rhs = st_psi + sizeof(C) * RB * EX;
    static constexpr size_t st_a = st_rhs + sizeof(C) * RB * EX;
    static constexpr size_t st_b = st_a + sizeof(R) * RB * EX;
    static constexpr size_t st_eps = st_b + sizeof(R) * RB * EX;
    static constexpr size_t st_nf = st_eps + (EPS ? sizeof(R) * RB * EX : 0);
    static constexpr size_t st_size = ((st_nf + RB * NFW + 127) / 128) * 128;
    static constexpr uint32_t st_tx_bytes = (uint32_t)(2 * sizeof(C) * RB * EX + (EPS ? 3 : 2) * sizeof(R) * RB * EX + RB * NFW);
    static constexpr size_t off_stage = 0;
    static constexpr size_t off_q = off_stage + NS * st_size;
    static constexpr size_t off_la = off_q + sizeof(C) * CD * W;
    static constexpr size_t off_lb = off_la + sizeof(C) * CD * W;
    static constexpr size_t off_dinv = off_lb + sizeof(C) * CD * W;
    static constexpr size_t off_ring = ((off_dinv + sizeof(R) * CD * W + 15) / 16) * 16;
    static constexpr size_t off_bar = ((off_ring + sizeof(C) * K * PS_RD * W + 15) / 16) * 16;
    static constexpr size_t total = off_bar + 8 * NS + 16;
};

// ------------------------------------------------------------------------------- the kernel
template <typename R, int K, int EX, int RB, int NS, bool TMA, bool EPS>
__global__ void __launch_bounds__(EX)
k_psi_stream(const __grid_constant__ StreamArgs A, const __grid_constant__ CUtensorMap tm_psi,
             const __grid_constant__ CUtensorMap tm_rhs, const __grid_constant__ CUtensorMap tm_a,
             const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_eps,
             const __grid_constant__ CUtensorMap tm_nf) {
    typedef typename V2<R>::type C;
    typedef StreamSmem<R, K, EX, RB, NS, EPS> S;
    constexpr int CD = S::CD, W = S::W, H = S::H, NFW = S::NFW, TX = EX - 2 * H;
    // declared alignment keeps the pointers in the shared address space (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(128) unsigned char smem[];
    C *q = (C *)(smem + S::off_q);
    C *la = (C *)(smem + S::off_la);
    C *lb = (C *)(smem + S::off_lb);
    R *dinv = (R *)(smem + S::off_dinv);
    C *ring = (C *)(smem + S::off_ring);
    uint64_t *full = (uint64_t *)(smem + S::off_bar);

    const Geo &g = A.g;
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * TX;
    const int x = x0 - H + t;
    const int xs16 = ((x0 - H + 1024) / 16) * 16 - 1024;   // 16-aligned start of the flag box (x0 - H >= -H)
    const int nfd = (x0 - H) - xs16;
    const int y0 = g.j0 + blockIdx.y * A.TY;
    const int y1 = min(y0 + A.TY, g.j1);
    const int yb = y0 - K;                  // first input row
    const int nin = (y1 - y0) + 2 * K;      // number of input rows
    const int nsteps = nin + K;             // last level lags the input by 2K rows and stops K rows early
    const bool xin = (x >= 0 && x < g.Nx);
    const bool xout = (t >= H && t < EX - H && x < g.Nx);
    const R dt = (R)A.dt, dx = (R)g.dx, dy = (R)g.dy, idx2 = (R)g.idx2, idy2 = (R)g.idy2;
    const R eps0 = (R)A.eps, lang = (R)A.lang_c;
    const R cx = dt * idx2, cy = dt * idy2;
    const int nchunks = (nin + RB - 1) / RB;

    auto stage_ptr = [&](int s) { return smem + S::off_stage + (size_t)s * S::st_size; };

    // ---- staging producers
    auto issue_chunk = [&](int chunk) {      // rows [chunk*RB, chunk*RB+RB) relative to yb
        int s = chunk % NS;
        unsigned char *st = stage_ptr(s);
        int prow = yb + chunk * RB - g.rb;   // plane row
        if (TMA) {
            if (t == 0) {
                uint32_t bytes = S::st_tx_bytes;
                if (A.same_rhs) bytes -= (uint32_t)(sizeof(C) * RB * EX);
                mbar_expect_tx(&full[s], bytes);
                const int cmul = sizeof(C) / 8 == 2 ? 2 : 1;     // complex double = two 8-byte elements
                tma_load_2d(st + S::st_psi, &tm_psi, (x0 - H) * cmul, prow, &full[s]);
                if (!A.same_rhs) tma_load_2d(st + S::st_rhs, &tm_rhs, (x0 - H) * cmul, prow, &full[s]);
                tma_load_2d(st + S::st_a, &tm_a, x0 - H, prow, &full[s]);
                tma_load_2d(st + S::st_b, &tm_b, x0 - H, prow, &full[s]);
                if (EPS) tma_load_2d(st + S::st_eps, &tm_eps, x0 - H, prow, &full[s]);
                tma_load_2d(st + S::st_nf, &tm_nf, xs16, prow, &full[s]);
            }
        } else {
            const C *gpsi = (const C *)A.psi, *grhs = (const C *)A.rhs;
            const R *ga = (const R *)A.a, *gb = (const R *)A.b, *ge = (const R *)A.epsf;
#pragma unroll
            for (int r = 0; r < RB; r++) {
                int pr = prow + r;
                bool ok = xin && pr >= 0 && pr < g.rows;
                size_t n = (size_t)(ok ? pr : 0) * g.P + (ok ? x : 0);
                C z; z.x = 0; z.y = 0;
                ((C *)(st + S::st_psi))[r * EX + t] = ok ? gpsi[n] : z;
                if (!A.same_rhs) ((C *)(st + S::st_rhs))[r * EX + t] = ok ? grhs[n] : z;
                ((R *)(st + S::st_a))[r * EX + t] = ok ? ga[n] : (R)0;
                ((R *)(st + S::st_b))[r * EX + t] = ok ? gb[n] : (R)0;
                if (EPS) ((R *)(st + S::st_eps))[r * EX + t] = ok ? ge[n] : (R)0;
                (st + S::st_nf)[r * NFW + nfd + t] = ok ? A.nf[n] : (uint8_t)0;
            }
        }
    };

    if (TMA && t == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    for (int c = 0; c < NS && c < nchunks; c++) issue_chunk(c);
    if (!TMA) __syncthreads();

    R rmax[K];
#pragma unroll
    for (int k = 0; k < K; k++) rmax[k] = 0;

    int c0 = 0;                              // step mod CD, kept incrementally
    int chunk = 0, ri = 0;                   // step / RB, step mod RB
    for (int step = 0; step < nsteps; step++) {
        // ---- arrival of input row `step`: constants + level-0 values
        if (step < nin) {
            int s = chunk % NS;
            if (TMA && ri == 0) mbar_wait(&full[s], (uint32_t)((chunk / NS) & 1));
            unsigned char *st = stage_ptr(s);
            C p0 = ((const C *)(st + S::st_psi))[ri * EX + t];
            C qq = A.same_rhs ? p0 : ((const C *)(st + S::st_rhs))[ri * EX + t];
            R av = ((const R *)(st + S::st_a))[ri * EX + t];
            R bv = ((const R *)(st + S::st_b))[ri * EX + t];
            unsigned f = (st + S::st_nf)[ri * NFW + nfd + t];
            R e = EPS ? ((const R *)(st + S::st_eps))[ri * EX + t] : eps0;
            if (!xin) f = 0;
            C La, Lb;
            La.x = La.y = Lb.x = Lb.y = 0;
            R di = 0;
            if (f) {
                if (A.noise) {
                    uint32_t nn = (uint32_t)x + (uint32_t)g.Nx * (uint32_t)(yb + step);
                    qq.x += lang * (rand_1<R>(nn, A.rand_t) - (R)0.5);
                    qq.y += lang * (rand_2<R>(nn, A.rand_t) - (R)0.5);
                }
                R sn, cs;
                if (f & (NF_PM | NF_PP)) { sincos_r<R>(dx * av, &sn, &cs); La.x = cx * cs; La.y = cx * sn; }
                if (f & (NF_MP | NF_PP)) { sincos_r<R>(dy * bv, &sn, &cs); Lb.x = cy * cs; Lb.y = cy * sn; }
                int nwx = ((f & (NF_MM | NF_MP)) ? 1 : 0) + ((f & (NF_PM | NF_PP)) ? 1 : 0);
                int nwy = ((f & (NF_MM | NF_PM)) ? 1 : 0) + ((f & (NF_MP | NF_PP)) ? 1 : 0);
                R D = (R)1.0 + dt * (qq.x * qq.x + qq.y * qq.y - e + (idx2 * (R)nwx + idy2 * (R)nwy));
                di = rcp_r(D);
            } else {
                qq.x = 0; qq.y = 0;
            }
            int cs_ = c0;
            q[cs_ * W + t + 1] = qq;
            la[cs_ * W + t + 1] = La;
            lb[cs_ * W + t + 1] = Lb;
            dinv[cs_ * W + t + 1] = di;
            ring[(0 * PS_RD + (step & (PS_RD - 1))) * W + t + 1] = p0;
        }
        // ---- levels: sweep k produces relative row step - 2k
#pragma unroll
        for (int k = 1; k <= K; k++) {
            int rk = step - 2 * k;
            if (rk >= k && rk <= nin - 1 - k) {
                const C *src = ring + (size_t)(k - 1) * PS_RD * W;
                int cs_ = c0 - 2 * k;          // (step - 2k) mod CD
                if (cs_ < 0) cs_ += CD;
                int csm = cs_ == 0 ? CD - 1 : cs_ - 1;
                int ps = (rk & (PS_RD - 1)) * W + t + 1;
                int pm = ((rk - 1) & (PS_RD - 1)) * W + t + 1, pp = ((rk + 1) & (PS_RD - 1)) * W + t + 1;
                C pw = src[ps - 1], pe = src[ps + 1], pS = src[pm], pN = src[pp], pc = src[ps];
                C lw = la[cs_ * W + t], le = la[cs_ * W + t + 1];
                C ls = lb[csm * W + t + 1], ln = lb[cs_ * W + t + 1];
                C acc = q[cs_ * W + t + 1];
                R di = dinv[cs_ * W + t + 1];
                // W: (c + i s) psi_W ; E: (c - i s) psi_E ; S: (c + i s) psi_S ; N: (c - i s) psi_N
                acc.x += lw.x * pw.x - lw.y * pw.y; acc.y += lw.x * pw.y + lw.y * pw.x;
                acc.x += le.x * pe.x + le.y * pe.y; acc.y += le.x * pe.y - le.y * pe.x;
                acc.x += ls.x * pS.x - ls.y * pS.y; acc.y += ls.x * pS.y + ls.y * pS.x;
                acc.x += ln.x * pN.x + ln.y * pN.y; acc.y += ln.x * pN.y - ln.y * pN.x;
                C nx;
                nx.x = acc.x * di; nx.y = acc.y * di;
                int row = yb + rk;
                bool oreg = xout && row >= y0 && row < y1;
                if (k < K) ring[((size_t)k * PS_RD + (rk & (PS_RD - 1))) * W + t + 1] = nx;
                else if (oreg) ((C *)A.out)[g.at(x, row)] = nx;
                if (oreg) rmax[k - 1] = fmax(rmax[k - 1], fmax(fabs(nx.x - pc.x), fabs(nx.y - pc.y)));
            }
        }
        __syncthreads();
        if (++c0 == CD) c0 = 0;
        // ---- refill the stage that was just drained
        if (++ri == RB) {
            ri = 0;
            int next = chunk + NS;
            if (next < nchunks) issue_chunk(next);
            chunk++;
        }
    }
    // ---- per-sweep max-norm updates -> one atomicMax per CTA and sweep
    __shared__ double sm_max[K][32];
    const int lane = t & 31, w = t >> 5, nw = EX / 32;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double r = warp_max((double)rmax[k]);
        if (lane == 0) sm_max[k][w] = r;
    }
    __syncthreads();
    if (t < K) {
        double r = 0.0;
        for (int i = 0; i < nw; i++) r = fmax(r, sm_max[t][i]);
        if (r > 0.0) atomicMax(A.slots + t, (unsigned long long)__double_as_longlong(r));
    }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
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

// 2-D map over a pitched plane: `width` valid elements per row, `rows` rows, pitch in bytes.
static int make_map(CUtensorMap *tm, CUtensorMapDataType dt, int esize, const void *base, size_t width, size_t rows,
                    size_t pitch_bytes, int box_w, int box_h) {
    PFN_encodeTiled enc = get_encode();
    SVL_REQUIRE(enc, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, (void *)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        svl_set_error("cuTensorMapEncodeTiled failed: %d (esize %d width %zu rows %zu pitch %zu box %dx%d)", (int)r, esize,
                      width, rows, pitch_bytes, box_w, box_h);
        return 1;
    }
    return 0;
}

template <typename R, int K, int EX, int RB, int NS, bool TMA, bool EPS>
static int launch_stream_t(svl_ctx *c, StreamArgs &A) {
    typedef typename V2<R>::type C;
    typedef StreamSmem<R, K, EX, RB, NS, EPS> S;
    const Geo &g = c->g;
    const int TX = EX - 2 * S::H;
    int nstrips = (g.Nx + TX - 1) / TX;
    int rows = g.j1 - g.j0;
    auto kern = k_psi_stream<R, K, EX, RB, NS, TMA, EPS>;
    // one wave: (resident CTAs per SM) x (SM count) CTAs at most; segments of at least 32 rows
    size_t smem = S::total;
    static int occ = 0, nsm = 0;          // per instantiation
    if (!occ) {
        SVL_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SVL_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, EX, smem));
        SVL_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
        if (occ < 1) occ = 1;
    }
    int want = (occ * nsm) / nstrips;
    if (want < 1) want = 1;
    int TY = (rows + want - 1) / want;
    if (TY < 32) TY = 32;
    if (TY > rows) TY = rows;
    int nsegs = (rows + TY - 1) / TY;
    A.TY = TY;
    CUtensorMap tm[6];
    memset(tm, 0, sizeof(tm));
    if (TMA) {
        const bool dbl = sizeof(R) == 8;
        CUtensorMapDataType rt = dbl ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
        // complex float is moved as one 8-byte element, complex double as two 8-byte elements
        CUtensorMapDataType ct = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
        int cmul = dbl ? 2 : 1;
        size_t pr = (size_t)g.P * sizeof(R), pc = (size_t)g.P * sizeof(C);
        SVL_TRY(make_map(&tm[0], ct, 8, A.psi, (size_t)g.Nx * cmul, g.rows, pc, EX * cmul, RB));
        SVL_TRY(make_map(&tm[1], ct, 8, A.rhs, (size_t)g.Nx * cmul, g.rows, pc, EX * cmul, RB));
        SVL_TRY(make_map(&tm[2], rt, sizeof(R), A.a, g.Nx, g.rows, pr, EX, RB));
        SVL_TRY(make_map(&tm[3], rt, sizeof(R), A.b, g.Nx, g.rows, pr, EX, RB));
        SVL_TRY(make_map(&tm[4], rt, sizeof(R), EPS ? A.epsf : A.a, g.Nx, g.rows, pr, EX, RB));
        SVL_TRY(make_map(&tm[5], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, A.nf, g.Nx, g.rows, (size_t)g.P, S::NFW, RB));
    }
    dim3 grid(nstrips, nsegs);
    kern<<<grid, EX, smem, c->stream>>>(A, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5]);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return 0;
}

template <typename R, int RB, int NS, bool EPS>
static int launch_stream_k(svl_ctx *c, int K, StreamArgs &A, bool tma) {
    if (!tma) {
        if (K == 4) return launch_stream_t<R, 4, 128, RB, NS, false, EPS>(c, A);
        svl_set_error("psi_stream: the plain-load staging variant is built for K=4 only");
        return 2;
    }
    switch (K) {
        case 1: return launch_stream_t<R, 1, 128, RB, NS, true, EPS>(c, A);
        case 2: return launch_stream_t<R, 2, 128, RB, NS, true, EPS>(c, A);
        case 3: return launch_stream_t<R, 3, 128, RB, NS, true, EPS>(c, A);
        case 4: return launch_stream_t<R, 4, 128, RB, NS, true, EPS>(c, A);
        case 6: return launch_stream_t<R, 6, 128, RB, NS, true, EPS>(c, A);
        case 8: return launch_stream_t<R, 8, 128, RB, NS, true, EPS>(c, A);
    }
    svl_set_error("psi_stream: K=%d not instantiated (1,2,3,4,6,8)", K);
    return 2;
}

// largest instantiated K that does not exceed the request (the solve driver chops sweep runs)
int svl_psi_stream_fit_k(int K) {
    static const int ks[] = {8, 6, 4, 3, 2, 1};
    for (int k : ks) if (k <= K) return k;
    return 1;
}

int svl_launch_psi_stream(svl_ctx *c, int K, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                          const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c, uint32_t rand_t,
                          unsigned long long *resid_slots) {
    StreamArgs A;
    memset(&A, 0, sizeof(A));
    A.g = c->g;
    A.dt = dt; A.eps = eps; A.lang_c = lang_c; A.rand_t = rand_t;
    A.has_eps = epsf != nullptr;
    A.noise = lang_c > 1.0e-32 ? 1 : 0;
    A.same_rhs = rhs->p[0] == psi->p[0];
    A.psi = psi->p[0]; A.rhs = rhs->p[0]; A.a = ab->p[0]; A.b = ab->p[1];
    A.epsf = epsf ? epsf->p[0] : nullptr;
    A.nf = c->nf; A.out = out->p[0]; A.slots = resid_slots;
    bool tma = c->opt_tma != 0;
    if (c->rsize == 4) {
        if (epsf) return launch_stream_k<float, 2, 2, true>(c, K, A, tma);
        return launch_stream_k<float, 2, 2, false>(c, K, A, tma);
    }
    if (epsf) return launch_stream_k<double, 2, 2, true>(c, K, A, tma);
    return launch_stream_k<double, 2, 2, false>(c, K, A, tma);
}
