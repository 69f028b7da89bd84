// CPU SIMT shim -- TEST INFRASTRUCTURE ONLY (never linked into the product .so).
//
// Lets g++ compile the *unmodified* CUDA C++ kernel sources of the reference
// (read from /root/reference/svirl/cuda/*.h at build time by oracle/build_ref.py,
// %-templated and concatenated exactly as svirl/parallel/startup.py:41-63 does)
// and run them on host cores.  It provides just what those kernels use:
//   __global__/__device__/__shared__, threadIdx/blockIdx/blockDim/gridDim, warpSize,
//   __syncthreads, __shfl_down_sync, atomicMax(int*), abs/max overloads and the
//   pycuda::complex<T> name (-> std::complex<T>).
// Execution models:
//   FLAT  : kernels without barriers/shuffles -- one plain call per thread.
//   FIBER : kernels with __syncthreads/__shfl_*: every CUDA thread of a block is a
//           ucontext fiber, scheduled round-robin; barriers/shuffles yield until the
//           block/warp has arrived, so the reduction trees execute in exactly the
//           reference's order.
// Blocks are distributed over OpenMP threads (all per-block state is thread_local).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <cmath>
#include <complex>
#include <vector>
#include <functional>
#include <ucontext.h>
#include <omp.h>

#define __global__
#define __device__
#define __host__
#define __inline__ inline
#define __forceinline__ inline
#define __shared__ thread_local
#define __restrict__

namespace pycuda { template <class T> using complex = std::complex<T>; }
using std::conj;

struct simt_dim3 { unsigned x, y, z; };
static thread_local simt_dim3 threadIdx, blockIdx, blockDim, gridDim;
static const int warpSize = 32;

inline float  max(float a, float b)   { return fmaxf(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline double max(float a, double b)  { return fmax((double)a, b); }
inline double max(double a, float b)  { return fmax(a, (double)b); }
inline float  min(float a, float b)   { return fminf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }

inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ------------------------------------------------------------------ fiber engine
struct simt_block_state {
    int nthreads = 0, cur = -1;
    bool fiber_mode = false;
    std::vector<ucontext_t> ctx;
    std::vector<char *> stacks;
    std::vector<char> done;
    ucontext_t main_ctx;
    long bar_count = 0, bar_gen = 0;
    long warp_count[32], warp_gen[32];
    uint64_t xchg[1024];
    const std::function<void()> *fn = nullptr;
};
static thread_local simt_block_state *simt_blk = nullptr;
static const size_t SIMT_STACK = 128 * 1024;

inline void simt_yield() {
    simt_block_state *b = simt_blk;
    if (!b || !b->fiber_mode) { fprintf(stderr, "simt_shim: barrier in FLAT mode\n"); abort(); }
    swapcontext(&b->ctx[b->cur], &b->main_ctx);
}

inline void simt_barrier(long &count, long &gen, int n) {
    long g = gen;
    if (++count == n) { count = 0; gen++; }
    else while (gen == g) simt_yield();
}

inline void __syncthreads() { simt_barrier(simt_blk->bar_count, simt_blk->bar_gen, simt_blk->nthreads); }

template <class T>
inline T __shfl_down_sync(unsigned, T v, int off) {
    simt_block_state *b = simt_blk;
    int t = threadIdx.x, w = t >> 5, l = t & 31;
    int wn = b->nthreads - w * 32; if (wn > 32) wn = 32;
    memcpy(&b->xchg[t], &v, sizeof(T));
    simt_barrier(b->warp_count[w], b->warp_gen[w], wn);
    T r = v;
    if (l + off < 32 && t + off < b->nthreads) memcpy(&r, &b->xchg[t + off], sizeof(T));
    simt_barrier(b->warp_count[w], b->warp_gen[w], wn);
    return r;
}

static void simt_fiber_entry() {
    simt_block_state *b = simt_blk;
    (*b->fn)();
    b->done[b->cur] = 1;
    swapcontext(&b->ctx[b->cur], &b->main_ctx);
}

inline void simt_run_block_fibers(int bx, const std::function<void()> &f) {
    if (!simt_blk) simt_blk = new simt_block_state();
    simt_block_state *b = simt_blk;
    b->fiber_mode = true; b->nthreads = bx; b->fn = &f;
    b->bar_count = 0; b->bar_gen = 0;
    for (int w = 0; w < 32; w++) { b->warp_count[w] = 0; b->warp_gen[w] = 0; }
    if ((int)b->ctx.size() < bx) {
        b->ctx.resize(bx);
        while ((int)b->stacks.size() < bx) b->stacks.push_back((char *)malloc(SIMT_STACK));
    }
    b->done.assign(bx, 0);
    for (int t = 0; t < bx; t++) {
        getcontext(&b->ctx[t]);
        b->ctx[t].uc_stack.ss_sp = b->stacks[t];
        b->ctx[t].uc_stack.ss_size = SIMT_STACK;
        b->ctx[t].uc_link = &b->main_ctx;
        makecontext(&b->ctx[t], simt_fiber_entry, 0);
    }
    int remaining = bx;
    while (remaining > 0) {
        for (int t = 0; t < bx; t++) {
            if (b->done[t]) continue;
            b->cur = t; threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            swapcontext(&b->main_ctx, &b->ctx[t]);
            if (b->done[t]) remaining--;
        }
    }
    b->fiber_mode = false;
}

enum { SIMT_FLAT = 0, SIMT_FIBER = 1 };

template <class F>
inline void simt_launch(int gx, int gy, int bx, int mode, F f) {
    const std::function<void()> fn = f;
    long nblocks = (long)gx * gy;
    #pragma omp parallel for schedule(static) if (nblocks >= 64)
    for (long blk = 0; blk < nblocks; blk++) {
        blockIdx.x = (unsigned)(blk % gx); blockIdx.y = (unsigned)(blk / gx); blockIdx.z = 0;
        blockDim.x = bx; blockDim.y = 1; blockDim.z = 1;
        gridDim.x = gx; gridDim.y = gy; gridDim.z = 1;
        if (mode == SIMT_FLAT) {
            for (int t = 0; t < bx; t++) { threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0; f(); }
        } else {
            simt_run_block_fibers(bx, fn);
        }
    }
}
