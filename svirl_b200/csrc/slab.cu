// Row-slab decomposition across the GPUs of one NVLink/NVSwitch box (new; the reference is
// single GPU: svirl/storage/arrays.py:16-17).  One process per GPU; rank r owns node rows
// [j0, j1) and every plane carries SVL_HALO rows of its neighbours.
//
//   * halo rows travel by DIRECT PEER STORES: after each launch a small kernel copies the
//     boundary rows of the buffer just written straight into the neighbours' planes (pointers
//     obtained through CUDA IPC), then publishes an epoch number in the neighbour's flag word
//     (st.release.sys after __threadfence_system);
//   * before the next launch a one-thread kernel spins (bounded) until both neighbours' epochs
//     have arrived (ld.acquire.sys).  By induction a rank can run at most one launch ahead of
//     its neighbours, which also makes the write-after-read on the rotating buffers safe;
//   * the only collective is the MAX of the per-sweep residual slots (exact, so the TD
//     trajectory is bitwise independent of the slab count); it goes through a host callback
//     (torch.distributed all_reduce behind it) once per solve batch.
#include "common.cuh"

// All exchanged planes are moved into ONE allocation (the "arena") so that a single CUDA IPC
// handle describes them: cudaMalloc carves small planes out of shared driver allocations, and an
// IPC handle always names a whole allocation.
struct SlabHandle {
    cudaIpcMemHandle_t h;          // 64 bytes
    unsigned long long off[10];    // byte offsets of psi x3, a x3, b x3, flags inside the arena
};

static int phys_of(const svl_ctx *c, const void *p) {
    for (int k = 0; k < 9; k++)
        if (c->own_phys[k] == p) return k;
    return -1;
}

// One kernel per push: blockIdx.y = direction (0: to the lower neighbour, 1: to the upper one).
// Every CTA copies its share of the boundary rows with 16-byte peer stores; the last CTA of a
// direction to finish publishes the epoch in the neighbour's flag word.
struct PushDesc {
    const unsigned char *src[2];
    unsigned char *dst[2];
    unsigned long long *flag[2];
    size_t bytes;
};

__global__ void __launch_bounds__(256)
k_push(PushDesc d, unsigned long long epoch, unsigned int *count) {
    const int dir = blockIdx.y;
    if (!d.dst[dir]) return;
    const uint4 *s = (const uint4 *)d.src[dir];
    uint4 *t = (uint4 *)d.dst[dir];
    size_t n = d.bytes / 16;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) t[i] = s[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(&count[dir], 1u);
        if (done == gridDim.x - 1) {
            count[dir] = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(d.flag[dir]), "l"(epoch) : "memory");
        }
    }
}

__global__ void k_wait_flags(const unsigned long long *flags, int has_lo, int has_hi, unsigned long long epoch, SpinGuard sg) {
    for (int s = 0; s < 2; s++) {
        if (!(s == 0 ? has_lo : has_hi)) continue;
        svl_spin_ge(flags + s, epoch, sg);
    }
}

// planes[] : 1 (psi) or 2 (a, b) planes of one buffer; they are pushed with ONE epoch
static int push_planes(svl_ctx *c, const void *const *planes, int nplanes, int esize, unsigned long long epoch) {
    const Geo &g = c->g;
    size_t pitch = (size_t)g.P * esize;
    int rows = g.j1 - g.j0;
    int depth = SVL_HALO < rows ? SVL_HALO : rows;
    for (int q = 0; q < nplanes; q++) {
        int id = phys_of(c, planes[q]);
        SVL_REQUIRE(id >= 0, "slab push: buffer is not one of the registered planes");
        PushDesc d;
        memset(&d, 0, sizeof(d));
        d.bytes = (size_t)depth * pitch;
        if (c->has_lo) {   // my lowest rows -> lower neighbour's upper halo (same global rows)
            d.src[0] = (const unsigned char *)planes[q] + (size_t)(g.j0 - g.rb) * pitch;
            d.dst[0] = (unsigned char *)c->peer[0][id] + (size_t)(g.j0 - c->nb_rb[0]) * pitch;
            d.flag[0] = c->peer_flags[0] + 1;          // I am the lower neighbour's "hi"
        }
        if (c->has_hi) {
            d.src[1] = (const unsigned char *)planes[q] + (size_t)(g.j1 - depth - g.rb) * pitch;
            d.dst[1] = (unsigned char *)c->peer[1][id] + (size_t)(g.j1 - depth - c->nb_rb[1]) * pitch;
            d.flag[1] = c->peer_flags[1] + 0;
        }
        // only the last plane of a buffer publishes the new epoch (stream order: earlier planes are complete)
        unsigned long long e = q == nplanes - 1 ? epoch : 0;
        if (q < nplanes - 1) { d.flag[0] = d.flag[0] ? c->scratch_flag : nullptr; d.flag[1] = d.flag[1] ? c->scratch_flag : nullptr; }
        int nb = (int)((d.bytes / 16 + 2047) / 2048);
        if (nb < 1) nb = 1;
        if (nb > 64) nb = 64;
        k_push<<<dim3(nb, 2), 256, 0, c->stream>>>(d, e, c->push_count);
        SVL_CHECK(cudaGetLastError());
    }
    return 0;
}

int svl_slab_push_psi(svl_ctx *c, const svl_buf *buf) {
    if (!c->slab_on) return 0;
    c->epoch_psi += 1;
    const void *pl[1] = {buf->p[0]};
    return push_planes(c, pl, 1, buf->esize, c->epoch_psi + c->epoch_A);
}

int svl_slab_push_ab(svl_ctx *c, const svl_buf *buf) {
    if (!c->slab_on) return 0;
    c->epoch_A += 1;
    const void *pl[2] = {buf->p[0], buf->p[1]};
    return push_planes(c, pl, 2, buf->esize, c->epoch_psi + c->epoch_A);
}

// The push of `out` will be done by the kernel that writes it: count it and describe the targets.
int svl_slab_push_fused(svl_ctx *c, const svl_buf *out, SlabPush *info) {
    memset(info, 0, sizeof(*info));
    if (!c->slab_on) return 0;
    const int nplanes = out->kind == SVL_EDGE ? 2 : 1;
    if (out->kind == SVL_EDGE) c->epoch_A += 1; else c->epoch_psi += 1;
    const Geo &g = c->g;
    int rows = g.j1 - g.j0;
    info->depth = SVL_HALO < rows ? SVL_HALO : rows;
    info->epoch = c->epoch_psi + c->epoch_A;
    info->count = c->push_count;
    for (int q = 0; q < nplanes; q++) {
        int id = phys_of(c, out->p[q]);
        SVL_REQUIRE(id >= 0, "slab push: buffer is not one of the registered planes");
        if (c->has_lo) info->peer[0][q] = c->peer[0][id];
        if (c->has_hi) info->peer[1][q] = c->peer[1][id];
    }
    info->peer_rb[0] = c->nb_rb[0]; info->peer_rb[1] = c->nb_rb[1];
    if (c->has_lo) info->flag[0] = c->peer_flags[0] + 1;      // I am the lower neighbour's "hi"
    if (c->has_hi) info->flag[1] = c->peer_flags[1] + 0;
    return 0;
}

unsigned long long svl_slab_epoch(svl_ctx *c) { return c->epoch_psi + c->epoch_A; }
void svl_slab_mark_waited(svl_ctx *c) { c->waited = c->epoch_psi + c->epoch_A; }

int svl_slab_wait(svl_ctx *c) {
    if (!c->slab_on || c->opt_slab_nocomm) return 0;
    unsigned long long e = c->epoch_psi + c->epoch_A;
    if (e == c->waited) return 0;
    k_wait_flags<<<1, 1, 0, c->stream>>>(c->flags, c->has_lo, c->has_hi, e, svl_spin_guard(c));
    SVL_CHECK(cudaGetLastError());
    c->waited = e;
    return 0;
}

// ----------------------------------------------------------------------------- setup ABI
// handles_out: one SlabHandle (144 bytes): IPC handle of the arena + offsets of the 9 planes and flags.
extern "C" int svl_slab_export(svl_ctx *c, svl_buf *psi, svl_buf *ab, void *handles_out) {
    SVL_REQUIRE(c && psi && ab && handles_out, "null argument");
    SVL_REQUIRE(psi->kind == SVL_NODE_C && ab->kind == SVL_EDGE, "psi NODE_C and ab EDGE required");
    SVL_REQUIRE(!c->arena, "slab arena already exported");
    svl_buf *ps[2], *as[2];
    for (int k = 0; k < 2; k++) { SVL_TRY(svl_scratch_node(c, k, &ps[k])); SVL_TRY(svl_scratch_edge(c, k, &as[k])); }
    svl_buf *bufs[9] = {psi, ps[0], ps[1], ab, as[0], as[1], ab, as[0], as[1]};
    int part[9] = {0, 0, 0, 0, 0, 0, 1, 1, 1};
    SlabHandle *h = (SlabHandle *)handles_out;
    size_t off = 0;
    for (int k = 0; k < 9; k++) {
        SVL_REQUIRE(!bufs[k]->borrowed, "buffer already lives in an arena");
        h->off[k] = off;
        off += (bufs[k]->bytes[part[k]] + 255) / 256 * 256;
    }
    h->off[9] = off;
    off += 256;
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    SVL_CHECK(cudaMalloc(&c->arena, off));
    SVL_CHECK(cudaMemset(c->arena, 0, off));
    for (int k = 0; k < 9; k++) {      // move the planes into the arena
        void *dst = (char *)c->arena + h->off[k];
        SVL_CHECK(cudaMemcpy(dst, bufs[k]->p[part[k]], bufs[k]->bytes[part[k]], cudaMemcpyDeviceToDevice));
        SVL_CHECK(cudaFree(bufs[k]->p[part[k]]));
        bufs[k]->p[part[k]] = dst;
        c->own_phys[k] = dst;
    }
    for (int k = 0; k < 6; k++) bufs[k]->borrowed = 1;
    c->flags = (unsigned long long *)((char *)c->arena + h->off[9]);
    c->scratch_flag = c->flags + 8;                    // sink for the non-final planes of a push
    SVL_CHECK(cudaMalloc(&c->push_count, 2 * sizeof(unsigned int)));
    SVL_CHECK(cudaMemset(c->push_count, 0, 2 * sizeof(unsigned int)));
    SVL_CHECK(cudaIpcGetMemHandle(&h->h, c->arena));
    SVL_CHECK(cudaDeviceSynchronize());
    return 0;
}

// lo/hi: the SlabHandle exported by the lower / upper neighbour (NULL at the ends of the chain),
// lo_j0 / hi_j0: the first owned row of that neighbour.
extern "C" int svl_slab_connect(svl_ctx *c, const void *lo, int lo_j0, const void *hi, int hi_j0) {
    SVL_REQUIRE(c, "null context");
    const void *hs[2] = {lo, hi};
    int j0s[2] = {lo_j0, hi_j0};
    for (int s = 0; s < 2; s++) {
        if (!hs[s]) continue;
        const SlabHandle *h = (const SlabHandle *)hs[s];
        void *base = nullptr;
        SVL_CHECK(cudaIpcOpenMemHandle(&base, h->h, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_base[s] = base;
        for (int k = 0; k < 9; k++) c->peer[s][k] = (char *)base + h->off[k];
        c->peer_flags[s] = (unsigned long long *)((char *)base + h->off[9]);
        c->nb_rb[s] = j0s[s] - SVL_HALO;
    }
    c->has_lo = lo != nullptr; c->has_hi = hi != nullptr;
    c->epoch_psi = c->epoch_A = c->waited = 0;
    c->slab_on = (c->has_lo || c->has_hi) ? 1 : 0;
    return 0;
}

extern "C" int svl_set_reduce_callback(svl_ctx *c, void (*reduce_max_u64)(unsigned long long *, int)) {
    SVL_REQUIRE(c, "null context");
    c->reduce_max_u64 = reduce_max_u64;
    return 0;
}

// Device-side variant: the callback reduces n device words in place with work enqueued on the
// context's stream (svl_get_stream), so no extra host synchronisation is needed.
extern "C" int svl_set_reduce_callback_device(svl_ctx *c, void (*reduce_max_dev)(unsigned long long *, int)) {
    SVL_REQUIRE(c, "null context");
    c->reduce_max_dev = reduce_max_dev;
    return 0;
}

extern "C" void *svl_get_stream(svl_ctx *c) { return c ? (void *)c->stream : nullptr; }

// Fill the halo rows of a field from the neighbours (used once after the fields were set).
extern "C" int svl_slab_exchange(svl_ctx *c, svl_buf *buf) {
    SVL_REQUIRE(c && buf, "null argument");
    if (!c->slab_on) return 0;
    if (buf->kind == SVL_NODE_C) SVL_TRY(svl_slab_push_psi(c, buf));
    else if (buf->kind == SVL_EDGE) SVL_TRY(svl_slab_push_ab(c, buf));
    else { svl_set_error("slab exchange: only psi / ab buffers are registered"); return 2; }
    SVL_TRY(svl_slab_wait(c));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return svl_peer_error(c);
}

// ----------------------------------------------------------------------------- residual board
// MAX of the per-sweep residual slots over all ranks without NCCL or the host: every rank writes
// its slots into every rank's board with peer stores, publishes an epoch word per destination, waits
// until all ranks' epochs have arrived on its own board and takes the maximum.  One small kernel
// per read-back (NVLink write latency + polling, a few microseconds) instead of an NCCL all-reduce
// driven from a Python callback.  Boards are double buffered by epoch parity: a rank cannot be two
// exchanges ahead of another one, because each exchange waits for everybody's epoch.
#define BOARD_WORDS (2 * SVL_MAX_RANKS * SVL_MAX_SWEEPS)      // slots[parity][rank][sweep], then epochs[rank]

struct BoardArgs {
    unsigned long long *peer[SVL_MAX_RANKS];
    int rank, world;
    SpinGuard sg;
};

__global__ void __launch_bounds__(256)
k_board_allmax(BoardArgs B, unsigned long long *d_resid, int first, int count, unsigned long long epoch) {
    const int par = (int)(epoch & 1ull);
    const size_t mine = ((size_t)(par * SVL_MAX_RANKS + B.rank)) * SVL_MAX_SWEEPS + first;
    for (int r = 0; r < B.world; r++) {
        unsigned long long *dst = B.peer[r] + mine;
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = d_resid[first + i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < B.world) {
        unsigned long long *e = B.peer[threadIdx.x] + BOARD_WORDS + B.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(e), "l"(epoch) : "memory");
        const unsigned long long *w = B.peer[B.rank] + BOARD_WORDS + threadIdx.x;
        svl_spin_ge(w, epoch, B.sg);
    }
    __syncthreads();
    const unsigned long long *own = B.peer[B.rank] + (size_t)par * SVL_MAX_RANKS * SVL_MAX_SWEEPS + first;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        unsigned long long m = 0;
        for (int r = 0; r < B.world; r++) {
            unsigned long long v = __ldcv(own + (size_t)r * SVL_MAX_SWEEPS + i);     // written by peers: bypass L1
            m = v > m ? v : m;
        }
        d_resid[first + i] = m;
    }
}

// Same exchange for sums (CG on slabs): every rank adds the contributions in RANK ORDER, so all ranks get
// the same bits and the host line search stays in lockstep.
__global__ void __launch_bounds__(64)
k_board_allsum(BoardArgs B, double *vals, int count, unsigned long long epoch) {
    const int par = (int)(epoch & 1ull);
    const size_t mine = ((size_t)(par * SVL_MAX_RANKS + B.rank)) * SVL_MAX_SWEEPS;
    for (int r = 0; r < B.world; r++)
        for (int i = threadIdx.x; i < count; i += blockDim.x)
            B.peer[r][mine + i] = (unsigned long long)__double_as_longlong(vals[i]);
    __syncthreads();
    if ((int)threadIdx.x < B.world) {
        __threadfence_system();
        unsigned long long *e = B.peer[threadIdx.x] + BOARD_WORDS + B.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(e), "l"(epoch) : "memory");
        const unsigned long long *w = B.peer[B.rank] + BOARD_WORDS + threadIdx.x;
        svl_spin_ge(w, epoch, B.sg);
    }
    __syncthreads();
    const unsigned long long *own = B.peer[B.rank] + (size_t)par * SVL_MAX_RANKS * SVL_MAX_SWEEPS;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < B.world; r++) s += __longlong_as_double((long long)__ldcv(own + (size_t)r * SVL_MAX_SWEEPS + i));
        vals[i] = s;
    }
}

int svl_board_allsum(svl_ctx *c, double *dvals, int count) {
    SVL_REQUIRE(c->board_world > 1, "sums over slabs need the residual board (svl_slab_board_connect)");
    SVL_REQUIRE(count > 0 && count <= SVL_MAX_SWEEPS, "too many values");
    BoardArgs B;
    memset(&B, 0, sizeof(B));
    for (int r = 0; r < c->board_world; r++) B.peer[r] = c->board_peer[r];
    B.rank = c->board_rank; B.world = c->board_world; B.sg = svl_spin_guard(c);
    c->board_epoch += 1;
    k_board_allsum<<<1, 64, 0, c->stream>>>(B, dvals, count, c->board_epoch);
    SVL_CHECK(cudaGetLastError());
    return 0;
}

int svl_board_allmax(svl_ctx *c, int first, int count) {
    BoardArgs B;
    memset(&B, 0, sizeof(B));
    for (int r = 0; r < c->board_world; r++) B.peer[r] = c->board_peer[r];
    B.rank = c->board_rank; B.world = c->board_world; B.sg = svl_spin_guard(c);
    c->board_epoch += 1;
    k_board_allmax<<<1, 256, 0, c->stream>>>(B, c->d_resid, first, count, c->board_epoch);
    SVL_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int svl_slab_board_export(svl_ctx *c, void *handle_out) {
    SVL_REQUIRE(c && handle_out, "null argument");
    SVL_REQUIRE(!c->board, "residual board already exported");
    size_t bytes = (size_t)(BOARD_WORDS + SVL_MAX_RANKS) * sizeof(unsigned long long);
    // a separate allocation of >= 2 MB gets its own driver allocation, hence its own IPC handle
    if (bytes < (4u << 20)) bytes = 4u << 20;
    SVL_CHECK(cudaMalloc(&c->board, bytes));
    SVL_CHECK(cudaMemset(c->board, 0, bytes));
    SVL_CHECK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out, c->board));
    SVL_CHECK(cudaDeviceSynchronize());
    return 0;
}

// handles: world x 64 bytes in rank order (this rank's own entry is ignored)
extern "C" int svl_slab_board_connect(svl_ctx *c, int rank, int world, const void *handles) {
    SVL_REQUIRE(c && handles && c->board, "export the board first");
    SVL_REQUIRE(world >= 1 && world <= SVL_MAX_RANKS && rank >= 0 && rank < world, "bad rank/world");
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
    for (int r = 0; r < world; r++) {
        if (r == rank) { c->board_peer[r] = c->board; continue; }
        void *p = nullptr;
        SVL_CHECK(cudaIpcOpenMemHandle(&p, h[r], cudaIpcMemLazyEnablePeerAccess));
        c->board_peer[r] = (unsigned long long *)p;
    }
    c->board_rank = rank; c->board_world = world; c->board_epoch = 0;
    return 0;
}
