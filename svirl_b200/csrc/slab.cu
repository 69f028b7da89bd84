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

// copy `depth` rows of `width_bytes` each: src/dst are plane base pointers, rows given as plane rows
__global__ void __launch_bounds__(256)
k_push_rows(const unsigned char *src, int src_row, unsigned char *dst, int dst_row, int depth, size_t pitch_bytes) {
    const uint4 *s = (const uint4 *)(src + (size_t)src_row * pitch_bytes);
    uint4 *d = (uint4 *)(dst + (size_t)dst_row * pitch_bytes);
    size_t n = (size_t)depth * pitch_bytes / 16;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) d[i] = s[i];
}

__global__ void k_publish(unsigned long long *flag, unsigned long long epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
}

__global__ void k_wait_flags(const unsigned long long *flags, int has_lo, int has_hi, unsigned long long epoch) {
    for (int s = 0; s < 2; s++) {
        if (!(s == 0 ? has_lo : has_hi)) continue;
        unsigned long long v = 0;
        long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + s) : "memory");
            if (clock64() - t0 > 20000000000ll) __trap();      // ~10 s: a neighbour died
        } while (v < epoch);
    }
}

static int push_plane(svl_ctx *c, const void *plane, int esize) {
    const Geo &g = c->g;
    int id = phys_of(c, plane);
    SVL_REQUIRE(id >= 0, "slab push: buffer is not one of the registered planes");
    size_t pitch = (size_t)g.P * esize;
    int rows = g.j1 - g.j0;
    int depth = SVL_HALO < rows ? SVL_HALO : rows;
    int nb = (int)((depth * pitch / 16 + 255) / 256);
    if (nb > 296) nb = 296;
    if (c->has_lo)   // my lowest rows -> lower neighbour's upper halo (same global rows)
        k_push_rows<<<nb, 256, 0, c->stream>>>((const unsigned char *)plane, g.j0 - g.rb, (unsigned char *)c->peer[0][id],
                                               g.j0 - c->nb_rb[0], depth, pitch);
    if (c->has_hi)
        k_push_rows<<<nb, 256, 0, c->stream>>>((const unsigned char *)plane, g.j1 - depth - g.rb, (unsigned char *)c->peer[1][id],
                                               g.j1 - depth - c->nb_rb[1], depth, pitch);
    SVL_CHECK(cudaGetLastError());
    return 0;
}

static int publish(svl_ctx *c) {
    unsigned long long e = c->epoch_psi + c->epoch_A;
    if (c->has_lo) k_publish<<<1, 1, 0, c->stream>>>(c->peer_flags[0] + 1, e);   // I am the lower neighbour's "hi"
    if (c->has_hi) k_publish<<<1, 1, 0, c->stream>>>(c->peer_flags[1] + 0, e);
    SVL_CHECK(cudaGetLastError());
    return 0;
}

int svl_slab_push_psi(svl_ctx *c, const svl_buf *buf) {
    if (!c->slab_on) return 0;
    SVL_TRY(push_plane(c, buf->p[0], buf->esize));
    c->epoch_psi += 1;
    return publish(c);
}

int svl_slab_push_ab(svl_ctx *c, const svl_buf *buf) {
    if (!c->slab_on) return 0;
    SVL_TRY(push_plane(c, buf->p[0], buf->esize));
    SVL_TRY(push_plane(c, buf->p[1], buf->esize));
    c->epoch_A += 1;
    return publish(c);
}

int svl_slab_wait(svl_ctx *c) {
    if (!c->slab_on) return 0;
    unsigned long long e = c->epoch_psi + c->epoch_A;
    if (e == c->waited) return 0;
    k_wait_flags<<<1, 1, 0, c->stream>>>(c->flags, c->has_lo, c->has_hi, e);
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

// Fill the halo rows of a field from the neighbours (used once after the fields were set).
extern "C" int svl_slab_exchange(svl_ctx *c, svl_buf *buf) {
    SVL_REQUIRE(c && buf, "null argument");
    if (!c->slab_on) return 0;
    if (buf->kind == SVL_NODE_C) SVL_TRY(svl_slab_push_psi(c, buf));
    else if (buf->kind == SVL_EDGE) SVL_TRY(svl_slab_push_ab(c, buf));
    else { svl_set_error("slab exchange: only psi / ab buffers are registered"); return 2; }
    SVL_TRY(svl_slab_wait(c));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}
