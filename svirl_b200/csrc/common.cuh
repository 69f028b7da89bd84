// Shared declarations of the svirl_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/svirl_b200.h"

#define SVL_HALO 8            // halo rows kept above/below the owned rows of every plane (>= max fused sweeps)
#define SVL_MAX_SWEEPS 1024   // svirl/solvers/td.py:164, 274
#define SVL_MAX_RANKS 16      // GPUs of one NVLink box that can share a residual board

// ----------------------------------------------------------------------------- errors
void svl_set_error(const char *fmt, ...);
#define SVL_CHECK(call)                                                                        \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            svl_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
#define SVL_REQUIRE(cond, msg)                                              \
    do {                                                                    \
        if (!(cond)) {                                                      \
            svl_set_error("%s:%d: %s", __FILE__, __LINE__, msg);            \
            return 2;                                                       \
        }                                                                   \
    } while (0)
#define SVL_TRY(call)            \
    do {                         \
        int rc_ = (call);        \
        if (rc_) return rc_;     \
    } while (0)

// ----------------------------------------------------------------------------- geometry
// Passed by value to every kernel.  Planes are pitched: element (i, j) of any plane lives at
// (j - rb) * P + i, where rb = j0 - SVL_HALO is the global row stored in plane row 0.
struct Geo {
    int Nx, Ny;        // global node counts
    int j0, j1;        // owned node rows [j0, j1)
    int rb;            // global row of plane row 0
    int P;             // pitch in elements (multiple of 32)
    int rows;          // plane rows = j1 - j0 + 2*SVL_HALO
    double dx, dy, idx, idy, idx2, idy2, idxy;
    __host__ __device__ __forceinline__ size_t at(int i, int j) const { return (size_t)(j - rb) * P + i; }
};

template <typename R> struct V2;
template <> struct V2<float>  { typedef float2 type; };
template <> struct V2<double> { typedef double2 type; };
__device__ __forceinline__ float fma_r(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_r(double a, double b, double c) { return fma(a, b, c); }

// node flag bits (svirl/cuda/td.h:48-57)
#define NF_MM 1
#define NF_MP 2
#define NF_PM 4
#define NF_PP 8

struct svl_buf {
    svl_ctx *ctx;
    int kind;
    int esize;          // bytes per element
    size_t n;           // elements in the flat (reference) layout
    void *p[2];         // plane(s): [0] nodes/cells/a/flat, [1] b
    size_t bytes[2];
    int borrowed;       // planes live in the slab arena (slab.cu): svl_free must not cudaFree them
};

struct svl_ctx {
    int device;
    int rsize;                     // 4 or 8
    Geo g;
    cudaStream_t stream;
    cudaStream_t stream2;          // slabs: boundary-tile launches (high priority), forked from / joined to `stream`
    cudaEvent_t ev_fork, ev_join, ev_go;
    // per-node material flags (always present)
    uint8_t *nf;
    bool have_mt;
    // scratch for solvers (allocated lazily)
    svl_buf *psi_s[2];
    svl_buf *ab_s[2];
    svl_buf *cg_s_node, *cg_s_edge;   // slabs: scratch for the CG directions (psi_s / ab_s live in the exchanged arena)
    // reductions
    double *partials;              // device, capacity partial_cap doubles
    size_t partial_cap;
    double *d_result;              // device, 64 doubles
    double *h_result;              // pinned, 64 doubles
    unsigned long long *d_resid;   // device, SVL_MAX_SWEEPS slots (bit patterns of non-negative doubles)
    unsigned long long *h_resid;   // pinned
    unsigned int *d_counter;       // last-block-done counters
    // vortex candidates
    long long *d_cand; double *d_candv; unsigned long long *d_ncand; size_t cand_cap;
    // options / stats
    int opt_psi_kernel, opt_psi_k, opt_psi_links, opt_psi_shape, opt_psi_patch, opt_tma, opt_graphs, opt_a_kernel, opt_cg_fused, opt_resid_board, opt_slab_nocomm, opt_slab_split, opt_slab_bnd, opt_cg_slabs;
    int pred_psi, pred_A;          // sweep counts of the previous solve
    int pred_psi2, pred_A2;        // ... and of the one before (trend)
    double stat_launches, stat_replays, stat_psi_sweeps, stat_A_sweeps;
    cudaEvent_t ev[8];
    // ---- slab decomposition (multi-GPU): see slab.cu
    int slab_on;                   // 1 once svl_slab_connect succeeded
    int has_lo, has_hi;            // neighbours below (smaller j) / above
    int nb_rb[2];                  // rb of the lower / upper neighbour
    void *peer[2][9];              // neighbour plane pointers by physical id: psi x3, a x3, b x3
    void *own_phys[9];             // this rank's planes by physical id
    void *arena;                   // one allocation holding the 9 exchanged planes + flags (one IPC handle)
    unsigned long long *flags;     // [2] written by the neighbours
    unsigned long long *scratch_flag;    // sink for pushes that must not publish
    unsigned long long *peer_flags[2];   // neighbour's flags array (we write slot [1] of lower, [0] of upper)
    unsigned long long epoch_psi, epoch_A;   // pushes issued so far
    unsigned long long waited;               // epoch the last wait kernel covered
    // residual / sum reduction across ranks (host callback; torch.distributed behind it)
    void (*reduce_max_u64)(unsigned long long *vals, int n);      // host values
    void (*reduce_max_dev)(unsigned long long *dvals, int n);     // device values, enqueued on c->stream
    unsigned int *push_count;      // [2] completion counters of the push kernel
    // residual board: every rank's per-sweep slots, written by the owners with peer stores (slab.cu)
    unsigned long long *board;           // own board (separate allocation, own IPC handle)
    unsigned long long *board_peer[SVL_MAX_RANKS];   // everybody's board, [rank] = own
    int board_rank, board_world;         // world == 0: not connected
    unsigned long long board_epoch;
    // diagnostics: per-launch timestamps of the slab tile kernels (option "trace" = number of launches)
    unsigned long long *trace;
    int trace_n, trace_cap;
    // bounded spin waits on peers (option "spin_timeout_ms", default 0 = wait forever): on a timeout the
    // waiting kernel raises *d_err (host-mapped pinned word) and carries on; the host turns it into an error
    long long spin_limit;          // clock64 cycles, 0 = unbounded
    int *h_err, *d_err;
    void *tma_cache;               // tensor-map descriptors of this context's planes (psi_tile.cu)
    // pipelined psi solves (td.cu): two banks of residual slots, the device-side go word of a pre-issued launch
    unsigned long long *d_resid_base, *h_resid_base;   // 2 x SVL_MAX_SWEEPS each; d_resid / h_resid point at the current bank
    int resid_bank;
    int *d_go;                     // written by k_psi_gate, read by the pre-issued tile launch
    const int *spec_gate;          // non-null while a gated launch is being issued (svl_launch_psi_tile picks it up)
    int spec_issued, spec_K;       // the next psi solve's first launch (spec_K sweeps) is already in the stream
    int opt_pipeline;              // option "pipeline" (default 1)
    int opt_pdl;                   // option "pdl" (default 1): batches of one solve as programmatic dependent launches
    double stat_spec_hit, stat_spec_miss;
    void *ipc_base[2];             // neighbours' arenas as mapped by cudaIpcOpenMemHandle (closed by svl_destroy)
};

// what a kernel needs to bound a spin wait on a peer
struct SpinGuard {
    long long limit;               // clock64 cycles, 0 = wait forever
    int *err;                      // host-visible flag raised on a timeout
};
static inline SpinGuard svl_spin_guard(const svl_ctx *c) { SpinGuard s = {c->spin_limit, c->d_err}; return s; }

static inline int svl_nblocks(size_t n, int b) { return (int)((n + b - 1) / b); }

// td.cu
int svl_launch_psi_sweep(svl_ctx *c, double dt, double eps, const svl_buf *epsf, const svl_buf *ab,
                         const svl_buf *rhs, const svl_buf *psi, svl_buf *out, double lang_c,
                         uint32_t rand_t, unsigned long long *resid_slot);
int svl_launch_a_sweep(svl_ctx *c, double dt, double kappa2, double rho, double H, const svl_buf *psi,
                       const svl_buf *ph, const svl_buf *rhs, const svl_buf *ab, svl_buf *out,
                       double lang_c, uint32_t rand_t, int write_rhs, unsigned long long *resid_slot);
// psi_tile.cu
void svl_tma_forget(svl_ctx *c, const void *base);             // drop cached tensor maps of a plane (nullptr: all)
// reduce.cu
int svl_ensure_partials(svl_ctx *c, size_t n);
int svl_finish_sum(svl_ctx *c, int nblocks, int nv, double scale, double *out_host);  // partials[nblocks*nv] -> host
// slab.cu
int svl_slab_push_psi(svl_ctx *c, const svl_buf *buf);       // boundary rows of a psi buffer -> neighbours' halos
int svl_slab_push_ab(svl_ctx *c, const svl_buf *buf);
int svl_slab_wait(svl_ctx *c);                               // wait until all pushes so far have arrived
int svl_board_allmax(svl_ctx *c, int first, int count);      // d_resid[first..] <- MAX over ranks (peer memory)
int svl_board_allsum(svl_ctx *c, double *dvals, int count);  // dvals[0..count) <- SUM over ranks, rank order (peer memory)
// In-kernel push (tile kernels): the tiles that own the first / last `depth` rows store their results
// into the neighbours' halo rows as well and the last of them publishes the epoch.
struct SlabPush {
    void *peer[2][2];              // [direction lo/hi][plane]: neighbour's copy of the output plane(s), or null
    int peer_rb[2];                // rb of that neighbour (its plane row 0 is global row peer_rb)
    unsigned long long *flag[2];   // neighbour's epoch word for this direction
    unsigned long long epoch;      // value to publish
    unsigned int *count;           // [2] completion counters (self-resetting)
    int depth;                     // rows pushed on each side
};
int svl_slab_push_fused(svl_ctx *c, const svl_buf *out, SlabPush *info);   // accounts the push, fills info
unsigned long long svl_slab_epoch(svl_ctx *c);               // pushes issued so far (what a consumer must wait for)
void svl_slab_mark_waited(svl_ctx *c);                       // the next kernel waits by itself
// abi.cu
int svl_peer_error(svl_ctx *c);                              // non-zero (and svl_last_error set) if a bounded peer wait timed out
int svl_scratch_node(svl_ctx *c, int k, svl_buf **out);
int svl_scratch_edge(svl_ctx *c, int k, svl_buf **out);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// Spin until *p >= epoch (a word written by a peer GPU with st.release.sys).  No wall-clock trap: a rank may
// legitimately lag by any amount (host I/O, line search); with a limit set, a timeout raises the host-visible
// error word and returns false -- the context survives and the next host call reports the failure.
__device__ __forceinline__ bool svl_spin_ge(const unsigned long long *p, unsigned long long epoch, const SpinGuard &sg) {
    unsigned long long v = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= epoch) return true;
        if (sg.limit > 0 && clock64() - t0 > sg.limit) {
            if (sg.err) { *(volatile int *)sg.err = 1; __threadfence_system(); }
            return false;
        }
    }
}
template <typename R> __device__ __forceinline__ void sincos_r(R x, R *s, R *c);
// fp32: branch-free Cody-Waite reduction (3-term pi/2) + minimax polynomials on [-pi/4, pi/4]
// (max abs error 7.6e-8 ~ 1.3 ulp for |x| < 2e4, measured against double; same class as sincosf,
// a third of its instructions).  Larger arguments take libdevice's sincosf.
struct sc_f { float s, c; };
static __device__ __noinline__ sc_f sincosf_slow(float x) { sc_f r; sincosf(x, &r.s, &r.c); return r; }
__device__ __forceinline__ void sincos_fast32(float x, float *s, float *c) {
    float j = rintf(x * 0.636619772f);
    int q = __float2int_rn(j);
    float r = fmaf(j, -1.5707962513e+0f, x);
    r = fmaf(j, -7.5497894159e-08f, r);
    r = fmaf(j, -5.3903029534e-15f, r);
    float z = r * r;
    float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float sn = fmaf(sp * z, r, r);
    float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    float cs = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
    float s2 = (q & 1) ? cs : sn, c2 = (q & 1) ? sn : cs;
    *s = (q & 2) ? -s2 : s2;
    *c = ((q + 1) & 2) ? -c2 : c2;
}
template <> __device__ __forceinline__ void sincos_r<float>(float x, float *s, float *c) {
    if (fabsf(x) > 20000.0f) { sc_f r = sincosf_slow(x); *s = r.s; *c = r.c; return; }   // out of line, by value
    sincos_fast32(x, s, c);
}
// fp64: the library's own sincos.  Quadrant by the 1.5*2^52 magic-number rounding (no F2I/I2F
// conversions), 3-term Cody-Waite reduction of pi/2 with FMAs (good for |x| < 1e5), fdlibm's degree-13/-12
// minimax kernels on [-pi/4, pi/4] (public-domain coefficients, kept in constant memory so they are FMA
// operands instead of per-use immediates), signs flipped through the high word.  Max error <= 2 ulp
// (tests/test_gpu_parity.py::test_sincos_accuracy); libdevice beyond the range.  ncu on the kernels that
// evaluate link variables showed them issue-slot bound with libdevice's sincos at ~130 dynamic instructions
// per link; the branch-free core lets callers interleave many evaluations (a_tile.cu).
struct sc_d { double s, c; };
static __device__ __noinline__ sc_d sincos_slow(double x) { sc_d r; sincos(x, &r.s, &r.c); return r; }
__constant__ double SVL_SC[16] = {
    // pi/2 split in three doubles; 2/pi
    1.5707963267948966e+00, 6.1232339957367574e-17, 8.4784276603688985e-32, 6.3661977236758138e-01,
    // sin: r + r z (S1 + z (S2 + ... z S6))
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    // cos: 1 - z/2 + z z (C1 + z (C2 + ... z C6))
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};
// branch-free core, valid for |x| <= SVL_SC_LIMIT; callers that evaluate many link variables test
// the range once for all of them (sincos_many_ok) so that the polynomials of different links
// interleave (instruction-level parallelism instead of one dependent Horner chain at a time).
#define SVL_SC_LIMIT 1.0e5
__device__ __forceinline__ void sincos_fast(double x, double *s, double *c) {
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: rint() in the low mantissa bits
    double t = fma(x, SVL_SC[3], MAGIC);
    int q = __double2loint(t);
    double j = t - MAGIC;
    double r = fma(j, -SVL_SC[0], x);
    r = fma(j, -SVL_SC[1], r);
    r = fma(j, -SVL_SC[2], r);
    double z = r * r;
    double ps = fma(SVL_SC[9], z, SVL_SC[8]);
    ps = fma(ps, z, SVL_SC[7]); ps = fma(ps, z, SVL_SC[6]); ps = fma(ps, z, SVL_SC[5]); ps = fma(ps, z, SVL_SC[4]);
    double sn = fma(ps * z, r, r);
    double pc = fma(SVL_SC[15], z, SVL_SC[14]);
    pc = fma(pc, z, SVL_SC[13]); pc = fma(pc, z, SVL_SC[12]); pc = fma(pc, z, SVL_SC[11]); pc = fma(pc, z, SVL_SC[10]);
    double cs = fma(pc * z, z, fma(-0.5, z, 1.0));
    // quadrant: swap for odd q, flip signs through the high word (bit 1 of q / q+1 -> bit 31)
    double s2 = (q & 1) ? cs : sn, c2 = (q & 1) ? sn : cs;
    int hs = __double2hiint(s2) ^ ((q << 30) & 0x80000000), hc = __double2hiint(c2) ^ (((q + 1) << 30) & 0x80000000);
    *s = __hiloint2double(hs, __double2loint(s2));
    *c = __hiloint2double(hc, __double2loint(c2));
}
__device__ __forceinline__ void sincos_fast(float x, float *s, float *c) { sincos_fast32(x, s, c); }
__device__ __forceinline__ bool sincos_fast_ok(double x) { return fabs(x) <= SVL_SC_LIMIT; }
__device__ __forceinline__ bool sincos_fast_ok(float x) { return fabsf(x) <= 20000.0f; }
// out-of-line libdevice evaluation (any argument): the rare path of callers that test the range themselves
__device__ __forceinline__ void sincos_any(double x, double *s, double *c) { sc_d r = sincos_slow(x); *s = r.s; *c = r.c; }
__device__ __forceinline__ void sincos_any(float x, float *s, float *c) { sc_f r = sincosf_slow(x); *s = r.s; *c = r.c; }
template <> __device__ __forceinline__ void sincos_r<double>(double x, double *s, double *c) {
    if (fabs(x) > SVL_SC_LIMIT) { sc_d r = sincos_slow(x); *s = r.s; *c = r.c; return; }
    sincos_fast(x, s, c);
}
// 1/x for x of order 1 (the Jacobi diagonal): MUFU.RCP + one Newton step, ~1 ulp, no slow path
__device__ __forceinline__ float rcp_r(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ double rcp_r(double x) { return __drcp_rn(x); }

// Thomas Wang hash RNG of the reference (svirl/cuda/common.h:36-63): exact integer maths.
__device__ __forceinline__ uint32_t wang_hash(uint32_t s) {
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s = s ^ (s >> 4);
    s *= 0x27d4eb2du;
    s = s ^ (s >> 15);
    return s;
}
template <typename R> __device__ __forceinline__ R rand_1(uint32_t n, uint32_t t) {
    return (R)(0.00000000023283064365386962890625 * (double)(R)wang_hash(71u * n + 9887u * t));
}
template <typename R> __device__ __forceinline__ R rand_2(uint32_t n, uint32_t t) {
    return (R)(0.00000000023283064365386962890625 * (double)(R)wang_hash(73u * n + 9901u * t + 1u));
}

// Im(conj(p0) U(ph) p1)  (svirl/cuda/common.h:65-73)
template <typename R, typename C>
__device__ __forceinline__ R js_link(C p0, R ph, C p1) {
    R s, c;
    sincos_r<R>(ph, &s, &c);
    return (p0.x * p1.y - p0.y * p1.x) * c - (p0.x * p1.x + p0.y * p1.y) * s;
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block max of non-negative values -> one atomicMax per CTA on the bit pattern (exact, order-free).
__device__ __forceinline__ void block_max_to_slot(double r, unsigned long long *slot) {
    __shared__ double sm_max[32];
    int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    r = warp_max(r);
    if ((tid & 31) == 0) sm_max[tid >> 5] = r;
    __syncthreads();
    if (tid < 32) {
        r = (tid < (nt + 31) / 32) ? sm_max[tid] : 0.0;
        r = warp_max(r);
        if (tid == 0 && r > 0.0) atomicMax(slot, (unsigned long long)__double_as_longlong(r));
    }
}

// Block sum of NV doubles per thread -> partials[block * NV + k] (fixed order => deterministic).
template <int NV>
__device__ __forceinline__ void block_sum_to_partials(double (&v)[NV], double *partials, int block_id) {
    __shared__ double sm_sum[32 * NV];
    int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    int lane = tid & 31, w = tid >> 5, nw = (nt + 31) / 32;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double x = warp_sum(v[k]);
        if (lane == 0) sm_sum[w * NV + k] = x;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = (lane < nw) ? sm_sum[lane * NV + k] : 0.0;
            x = warp_sum(x);
            if (lane == 0) partials[(size_t)block_id * NV + k] = x;
        }
    }
}
#endif
