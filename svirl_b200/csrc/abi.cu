// Context, buffers, copies, options: the non-arithmetic part of the C ABI.
// Replaces svirl/parallel/startup.py (context), svirl/storage/arrays.py (device side),
// svirl/parallel/utils.py (copy_dtod) of the reference.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

void svl_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *svl_last_error(void) { return g_err; }
extern "C" int svl_version(void) { return 100; }

// ----------------------------------------------------------------------------- node flags
__global__ void k_node_flags(Geo g, const uint8_t *mt, uint8_t *nf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int jl = blockIdx.y * blockDim.y + threadIdx.y;   // plane row
    if (i >= g.P || jl >= g.rows) return;
    int j = jl + g.rb;
    uint8_t f = 0;
    if (i < g.Nx && j >= 0 && j < g.Ny) {
        bool mm = (i > 0) && (j > 0), mp = (i > 0) && (j + 1 < g.Ny);
        bool pm = (i + 1 < g.Nx) && (j > 0), pp = (i + 1 < g.Nx) && (j + 1 < g.Ny);
        if (mt) {   // cells are stored in a pitched plane too: cell (ci, cj) at g.at(ci, cj)
            // plane row 0 of a slab that does not start at the bottom of the grid: the cell row below it is not
            // stored (found by compute-sanitizer memcheck).  That row is the outermost halo ring, whose own update
            // is never used, so its two lower bits may be anything: take 0.
            const bool below = jl > 0;
            if (mm) mm = below && mt[g.at(i - 1, j - 1)] != 0;
            if (mp) mp = mt[g.at(i - 1, j)] != 0;
            if (pm) pm = below && mt[g.at(i, j - 1)] != 0;
            if (pp) pp = mt[g.at(i, j)] != 0;
        }
        f = (mm ? NF_MM : 0) | (mp ? NF_MP : 0) | (pm ? NF_PM : 0) | (pp ? NF_PP : 0);
    }
    nf[(size_t)jl * g.P + i] = f;
}

extern "C" int svl_set_material(svl_ctx *c, const svl_buf *mt) {
    SVL_REQUIRE(c, "null context");
    SVL_REQUIRE(!mt || mt->kind == SVL_CELL_B, "material tiling must be a SVL_CELL_B buffer");
    dim3 b(32, 8), gr((c->g.P + 31) / 32, (c->g.rows + 7) / 8);
    k_node_flags<<<gr, b, 0, c->stream>>>(c->g, mt ? (const uint8_t *)mt->p[0] : nullptr, c->nf);
    SVL_CHECK(cudaGetLastError());
    c->have_mt = mt != nullptr;
    return 0;
}

// ----------------------------------------------------------------------------- lifecycle
extern "C" int svl_create(svl_ctx **out, int device_id, int Nx, int Ny, double dx, double dy, int dtype_bytes,
                          int j0, int j1) {
    SVL_REQUIRE(out, "null out");
    SVL_REQUIRE(dtype_bytes == 4 || dtype_bytes == 8, "dtype_bytes must be 4 or 8");
    SVL_REQUIRE(Nx >= 4 && Ny >= 4, "Nx, Ny must be >= 4");
    SVL_REQUIRE(0 <= j0 && j0 < j1 && j1 <= Ny, "bad slab rows");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        svl_set_error("no CUDA device available (%s): svirl_b200 has no CPU fallback", cudaGetErrorString(e));
        return 3;
    }
    SVL_REQUIRE(device_id >= 0 && device_id < ndev, "bad device id");
    SVL_CHECK(cudaSetDevice(device_id));
    svl_ctx *c = new svl_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device_id;
    c->rsize = dtype_bytes;
    Geo &g = c->g;
    g.Nx = Nx; g.Ny = Ny; g.j0 = j0; g.j1 = j1;
    g.rb = j0 - SVL_HALO;
    g.rows = j1 - j0 + 2 * SVL_HALO;
    g.P = ((Nx + 31) / 32) * 32;
    g.dx = dx; g.dy = dy;
    g.idx = 1.0 / dx; g.idy = 1.0 / dy;
    g.idx2 = 1.0 / (dx * dx); g.idy2 = 1.0 / (dy * dy); g.idxy = 1.0 / (dx * dy);
    SVL_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        SVL_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SVL_CHECK(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
        SVL_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        SVL_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        SVL_CHECK(cudaEventCreateWithFlags(&c->ev_go, cudaEventDisableTiming));
    }
    size_t nfb = (size_t)g.rows * g.P;
    SVL_CHECK(cudaMalloc(&c->nf, nfb));
    SVL_CHECK(cudaMalloc(&c->d_result, 64 * sizeof(double)));
    SVL_CHECK(cudaMallocHost(&c->h_result, 64 * sizeof(double)));
    SVL_CHECK(cudaMalloc(&c->d_resid_base, 2 * SVL_MAX_SWEEPS * sizeof(unsigned long long)));
    SVL_CHECK(cudaMallocHost(&c->h_resid_base, 2 * SVL_MAX_SWEEPS * sizeof(unsigned long long)));
    c->d_resid = c->d_resid_base; c->h_resid = c->h_resid_base; c->resid_bank = 0;
    SVL_CHECK(cudaMalloc(&c->d_go, sizeof(int)));
    SVL_CHECK(cudaMemsetAsync(c->d_go, 0, sizeof(int), c->stream));
    c->opt_pipeline = 1;
    { const char *e = getenv("SVL_PDL"); c->opt_pdl = e ? atoi(e) : 2; }   // 0 off, 1 one GPU only, 2 also the launch pairs of slab batches
    SVL_CHECK(cudaMalloc(&c->d_counter, 16 * sizeof(unsigned int)));
    SVL_CHECK(cudaMemsetAsync(c->d_counter, 0, 16 * sizeof(unsigned int), c->stream));
    SVL_CHECK(cudaMalloc(&c->d_ncand, sizeof(unsigned long long)));
    for (int k = 0; k < 8; k++) SVL_CHECK(cudaEventCreate(&c->ev[k]));
    SVL_CHECK(cudaHostAlloc(&c->h_err, sizeof(int), cudaHostAllocMapped));
    *c->h_err = 0;
    SVL_CHECK(cudaHostGetDevicePointer(&c->d_err, c->h_err, 0));
    {   // bound of the spin waits on peer GPUs: off unless asked for (SVL_SPIN_TIMEOUT_MS or option "spin_timeout_ms")
        const char *e = getenv("SVL_SPIN_TIMEOUT_MS");
        c->spin_limit = e ? (long long)(atof(e) * 2.0e6) : 0;
    }
    c->opt_psi_kernel = 2;
    c->opt_psi_k = 4;
    {   // fp32 link variables of the tile kernel: MUFU sin/cos after an exact reduction (1, default) or the
        // polynomial sincos (0); see link_sincos in psi_tile.cu.  fp64 always uses the polynomial.
        const char *e = getenv("SVL_PSI_LINKS");
        c->opt_psi_links = e ? atoi(e) : 1;
    }
    { const char *e = getenv("SVL_PSI_PATCH"); c->opt_psi_patch = e ? atoi(e) : 1; }   // 2 x 4 patch kernel (bit-identical to the column kernel, 3.4 % faster at cfg2)
    c->opt_psi_shape = 1;          // 256 threads x 8 rows (measured 63.2 vs 65.1 us per K=4 launch at 2048^2 for 512 x 4)
    c->opt_tma = 1;
    c->opt_graphs = 1;
    c->opt_a_kernel = 2;
    c->opt_cg_fused = 2;
    c->opt_resid_board = 1;
    c->opt_slab_split = 1;
    c->pred_psi = c->pred_A = c->pred_psi2 = c->pred_A2 = 0;
    *out = c;
    return svl_set_material(c, nullptr);
}

extern "C" int svl_destroy(svl_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 2; k++) {
        if (c->psi_s[k]) svl_free(c, c->psi_s[k]);
        if (c->ab_s[k]) svl_free(c, c->ab_s[k]);
    }
    if (c->cg_s_node) svl_free(c, c->cg_s_node);
    if (c->cg_s_edge) svl_free(c, c->cg_s_edge);
    // peer mappings first: the neighbours' arenas and boards opened over CUDA IPC
    for (int s = 0; s < 2; s++)
        if (c->ipc_base[s]) cudaIpcCloseMemHandle(c->ipc_base[s]);
    for (int r = 0; r < c->board_world; r++)
        if (r != c->board_rank && c->board_peer[r]) cudaIpcCloseMemHandle(c->board_peer[r]);
    svl_tma_forget(c, nullptr);
    cudaFree(c->arena);
    cudaFree(c->board);
    cudaFree(c->nf); cudaFree(c->d_result); cudaFreeHost(c->h_result);
    cudaFree(c->d_resid_base); cudaFreeHost(c->h_resid_base); cudaFree(c->d_counter); cudaFree(c->d_go);
    cudaFree(c->partials); cudaFree(c->d_cand); cudaFree(c->d_candv); cudaFree(c->d_ncand);
    for (int k = 0; k < 8; k++) cudaEventDestroy(c->ev[k]);
    cudaFreeHost(c->h_err);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->stream2);
    cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); cudaEventDestroy(c->ev_go);
    delete c;
    return 0;
}

int svl_peer_error(svl_ctx *c) {
    if (c->h_err && *c->h_err) {
        *c->h_err = 0;
        svl_set_error("a spin wait on a peer GPU timed out (option spin_timeout_ms): a rank is late or gone");
        return 4;
    }
    return 0;
}

extern "C" int svl_synchronize(svl_ctx *c) {
    SVL_REQUIRE(c, "null context");
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return svl_peer_error(c);
}

// CUDA events on the stream every kernel of this context is launched on (torch.cuda.Event would
// only see torch's current stream).
extern "C" int svl_event_record(svl_ctx *c, int slot) {
    SVL_REQUIRE(c && slot >= 0 && slot < 8, "bad event slot");
    SVL_CHECK(cudaEventRecord(c->ev[slot], c->stream));
    return 0;
}

extern "C" int svl_event_elapsed_ms(svl_ctx *c, int slot0, int slot1, double *ms) {
    SVL_REQUIRE(c && ms && slot0 >= 0 && slot0 < 8 && slot1 >= 0 && slot1 < 8, "bad event slot");
    SVL_CHECK(cudaEventSynchronize(c->ev[slot1]));
    float f = 0.f;
    SVL_CHECK(cudaEventElapsedTime(&f, c->ev[slot0], c->ev[slot1]));
    *ms = (double)f;
    return 0;
}

extern "C" int svl_set_option(svl_ctx *c, const char *name, int v) {
    SVL_REQUIRE(c && name, "null argument");
    if (!strcmp(name, "psi_kernel")) c->opt_psi_kernel = v;
    else if (!strcmp(name, "psi_k")) { SVL_REQUIRE(v >= 1 && v <= SVL_HALO, "psi_k out of range"); c->opt_psi_k = v; }
    else if (!strcmp(name, "psi_links")) c->opt_psi_links = v;         // fp32 tile kernel: 1 = link variables by MUFU sin/cos (psi_tile.cu)
    else if (!strcmp(name, "pdl")) c->opt_pdl = v;
    else if (!strcmp(name, "pipeline")) c->opt_pipeline = v;           // kappa = inf time stepping: pre-issue the next step's first launch behind a device-side gate (td.cu)
    else if (!strcmp(name, "psi_patch")) c->opt_psi_patch = v;         // fp32 tile kernel: 1 = 2 x 4 node patch per thread (k_psi_patch), 0 = 1 x 8 column
    else if (!strcmp(name, "psi_shape")) c->opt_psi_shape = v;         // fp32 tile kernel: 0 = 512 threads x 4 rows, 1 = 256 threads x 8 rows
    else if (!strcmp(name, "tma")) c->opt_tma = v;
    else if (!strcmp(name, "graphs")) c->opt_graphs = v;
    else if (!strcmp(name, "spin_timeout_ms")) c->spin_limit = (long long)v * 2000000ll;   // ~2 GHz clock64 ticks; 0 = wait forever
    else if (!strcmp(name, "a_kernel")) c->opt_a_kernel = v;
    else if (!strcmp(name, "cg_fused")) c->opt_cg_fused = v;
    else if (!strcmp(name, "resid_board")) c->opt_resid_board = v;
    else if (!strcmp(name, "slab_split")) c->opt_slab_split = v;
    else if (!strcmp(name, "slab_bnd")) c->opt_slab_bnd = v;           // CTAs of the boundary launch: 0 = automatic, -1 = one per boundary tile (round-2a behaviour)
    else if (!strcmp(name, "cg_slabs")) c->opt_cg_slabs = v;           // no-op since round 2: CG / energy on row slabs is always on
    else if (!strcmp(name, "slab_nocomm")) c->opt_slab_nocomm = v;     // timing experiments only: ranks run uncoupled
    else if (!strcmp(name, "trace")) {                                 // diagnostics: record v launches from now on
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        cudaFree(c->trace); c->trace = nullptr; c->trace_n = c->trace_cap = 0;
        if (v > 0) {
            std::vector<unsigned long long> init((size_t)4 * v, 0ull);
            for (int k = 0; k < v; k++) init[4 * k] = ~0ull;
            SVL_CHECK(cudaMalloc(&c->trace, init.size() * sizeof(unsigned long long)));
            SVL_CHECK(cudaMemcpy(c->trace, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
            c->trace_cap = v;
        }
    }
    else if (!strcmp(name, "reset_prediction")) { c->pred_psi = c->pred_A = c->pred_psi2 = c->pred_A2 = 0; }
    else { svl_set_error("unknown option %s", name); return 2; }
    return 0;
}

extern "C" int svl_get_stat(svl_ctx *c, const char *name, double *v) {
    SVL_REQUIRE(c && name && v, "null argument");
    if (!strcmp(name, "launches")) *v = c->stat_launches;
    else if (!strcmp(name, "cg_fused")) *v = c->opt_cg_fused;
    else if (!strcmp(name, "slab_on")) *v = c->slab_on;
    else if (!strcmp(name, "replays")) *v = c->stat_replays;
    else if (!strcmp(name, "spec_hit")) *v = c->stat_spec_hit;
    else if (!strcmp(name, "spec_miss")) *v = c->stat_spec_miss;
    else if (!strcmp(name, "psi_sweeps")) *v = c->stat_psi_sweeps;
    else if (!strcmp(name, "A_sweeps")) *v = c->stat_A_sweeps;
    else if (!strcmp(name, "pitch")) *v = c->g.P;
    else if (!strcmp(name, "reset")) { c->stat_launches = c->stat_replays = c->stat_psi_sweeps = c->stat_A_sweeps = 0; *v = 0; }
    else { svl_set_error("unknown stat %s", name); return 2; }
    return 0;
}

// ----------------------------------------------------------------------------- buffers
static int kind_esize(const svl_ctx *c, int kind) {
    switch (kind) {
        case SVL_NODE_C: return 2 * c->rsize;
        case SVL_NODE_R: case SVL_EDGE: case SVL_CELL_R: return c->rsize;
        case SVL_CELL_B: return 1;
    }
    return 0;
}

// number of valid global rows / row width (elements) of plane `part` of a kind
static void plane_shape(const svl_ctx *c, int kind, int part, int *nrows, int *width) {
    const Geo &g = c->g;
    switch (kind) {
        case SVL_NODE_C: case SVL_NODE_R: *nrows = g.Ny; *width = g.Nx; break;
        case SVL_EDGE:
            if (part == 0) { *nrows = g.Ny; *width = g.Nx - 1; }
            else { *nrows = g.Ny - 1; *width = g.Nx; }
            break;
        default: *nrows = g.Ny - 1; *width = g.Nx - 1; break;
    }
}

extern "C" int svl_alloc(svl_ctx *c, int kind, size_t n, int elem_size, svl_buf **out) {
    SVL_REQUIRE(c && out, "null argument");
    SVL_REQUIRE(kind >= SVL_NODE_C && kind <= SVL_FLAT, "bad kind");
    SVL_CHECK(cudaSetDevice(c->device));
    svl_buf *b = new svl_buf();
    memset(b, 0, sizeof(*b));
    b->ctx = c; b->kind = kind;
    const Geo &g = c->g;
    if (kind == SVL_FLAT) {
        SVL_REQUIRE(elem_size > 0, "elem_size required for SVL_FLAT");
        b->esize = elem_size; b->n = n;
        b->bytes[0] = (n ? n : 1) * (size_t)elem_size;
    } else {
        b->esize = kind_esize(c, kind);
        size_t plane = (size_t)g.rows * g.P * b->esize;
        b->bytes[0] = plane;
        size_t N = (size_t)g.Nx * g.Ny;
        switch (kind) {
            case SVL_NODE_C: case SVL_NODE_R: b->n = N; break;
            case SVL_EDGE: b->n = (size_t)(g.Nx - 1) * g.Ny + (size_t)g.Nx * (g.Ny - 1); b->bytes[1] = plane; break;
            default: b->n = (size_t)(g.Nx - 1) * (g.Ny - 1); break;
        }
    }
    for (int k = 0; k < 2; k++) {
        if (!b->bytes[k]) continue;
        cudaError_t e = cudaMalloc(&b->p[k], b->bytes[k]);
        if (e != cudaSuccess) {
            svl_set_error("cudaMalloc of %zu bytes failed: %s", b->bytes[k], cudaGetErrorString(e));
            if (k == 1) cudaFree(b->p[0]);
            delete b;
            return 1;
        }
        SVL_CHECK(cudaMemsetAsync(b->p[k], 0, b->bytes[k], c->stream));
    }
    *out = b;
    return 0;
}

extern "C" int svl_free(svl_ctx *c, svl_buf *b) {
    if (!b) return 0;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    svl_tma_forget(c, b->p[0]);
    if (b->p[1]) svl_tma_forget(c, b->p[1]);
    if (!b->borrowed) { cudaFree(b->p[0]); cudaFree(b->p[1]); }
    delete b;
    return 0;
}

extern "C" size_t svl_buf_size(const svl_buf *b) { return b ? b->n : 0; }

// Copy global rows [r0, r1) of one plane between the flat host layout and the pitched plane.
static int copy_rows(svl_ctx *c, const svl_buf *b, int part, int r0, int r1, void *host, size_t host_row0,
                     bool to_device) {
    int nrows, width;
    plane_shape(c, b->kind, part, &nrows, &width);
    const Geo &g = c->g;
    int lo = r0 > g.rb ? r0 : g.rb, hi = r1 < g.rb + g.rows ? r1 : g.rb + g.rows;
    if (lo < 0) lo = 0;
    if (hi > nrows) hi = nrows;
    if (hi <= lo) return 0;
    size_t wbytes = (size_t)width * b->esize, pbytes = (size_t)g.P * b->esize;
    char *d = (char *)b->p[part] + (size_t)(lo - g.rb) * pbytes;
    char *h = (char *)host + ((size_t)lo - host_row0) * wbytes;
    if (to_device) SVL_CHECK(cudaMemcpy2DAsync(d, pbytes, h, wbytes, wbytes, hi - lo, cudaMemcpyHostToDevice, c->stream));
    else SVL_CHECK(cudaMemcpy2DAsync(h, wbytes, d, pbytes, wbytes, hi - lo, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

extern "C" int svl_h2d(svl_ctx *c, svl_buf *dst, const void *src) {
    SVL_REQUIRE(c && dst && src, "null argument");
    SVL_CHECK(cudaSetDevice(c->device));
    if (dst->kind == SVL_FLAT) {
        SVL_CHECK(cudaMemcpyAsync(dst->p[0], src, dst->n * dst->esize, cudaMemcpyHostToDevice, c->stream));
    } else {
        const Geo &g = c->g;
        // device gets owned rows plus whatever halo rows exist globally
        SVL_TRY(copy_rows(c, dst, 0, g.rb, g.rb + g.rows, (void *)src, 0, true));
        if (dst->kind == SVL_EDGE) {
            const char *sb = (const char *)src + (size_t)(g.Nx - 1) * g.Ny * dst->esize;
            SVL_TRY(copy_rows(c, dst, 1, g.rb, g.rb + g.rows, (void *)sb, 0, true));
        }
    }
    SVL_CHECK(cudaStreamSynchronize(c->stream));   // host buffer may be pageable / reused by the caller
    return 0;
}

extern "C" int svl_d2h(svl_ctx *c, void *dst, const svl_buf *src) {
    SVL_REQUIRE(c && dst && src, "null argument");
    SVL_CHECK(cudaSetDevice(c->device));
    if (src->kind == SVL_FLAT) {
        SVL_CHECK(cudaMemcpyAsync(dst, src->p[0], src->n * src->esize, cudaMemcpyDeviceToHost, c->stream));
    } else {
        const Geo &g = c->g;
        SVL_TRY(copy_rows(c, src, 0, g.j0, g.j1, dst, 0, false));   // owned rows only
        if (src->kind == SVL_EDGE) {
            char *db = (char *)dst + (size_t)(g.Nx - 1) * g.Ny * src->esize;
            SVL_TRY(copy_rows(c, src, 1, g.j0, g.j1, db, 0, false));
        }
    }
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int svl_h2d_rows(svl_ctx *c, svl_buf *dst, int part, int r0, int r1, const void *src) {
    SVL_REQUIRE(c && dst && src && dst->kind != SVL_FLAT, "bad argument");
    SVL_REQUIRE(part == 0 || (part == 1 && dst->kind == SVL_EDGE), "bad part");
    SVL_CHECK(cudaSetDevice(c->device));
    SVL_TRY(copy_rows(c, dst, part, r0, r1, (void *)src, (size_t)r0, true));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int svl_d2h_rows(svl_ctx *c, void *dst, const svl_buf *src, int part, int r0, int r1) {
    SVL_REQUIRE(c && dst && src && src->kind != SVL_FLAT, "bad argument");
    SVL_REQUIRE(part == 0 || (part == 1 && src->kind == SVL_EDGE), "bad part");
    SVL_CHECK(cudaSetDevice(c->device));
    SVL_TRY(copy_rows(c, src, part, r0, r1, dst, (size_t)r0, false));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int svl_d2d(svl_ctx *c, svl_buf *dst, const svl_buf *src) {
    SVL_REQUIRE(c && dst && src, "null argument");
    SVL_REQUIRE(dst->kind == src->kind && dst->bytes[0] == src->bytes[0] && dst->bytes[1] == src->bytes[1],
                "d2d: buffers differ in kind/size");
    for (int k = 0; k < 2; k++)
        if (src->bytes[k])
            SVL_CHECK(cudaMemcpyAsync(dst->p[k], src->p[k], src->bytes[k], cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

extern "C" int svl_fill_zero(svl_ctx *c, svl_buf *b) {
    SVL_REQUIRE(c && b, "null argument");
    for (int k = 0; k < 2; k++)
        if (b->bytes[k]) SVL_CHECK(cudaMemsetAsync(b->p[k], 0, b->bytes[k], c->stream));
    return 0;
}

extern "C" int svl_swap(svl_ctx *c, svl_buf *x, svl_buf *y) {
    SVL_REQUIRE(c && x && y, "null argument");
    SVL_REQUIRE(x->kind == y->kind && x->bytes[0] == y->bytes[0] && x->bytes[1] == y->bytes[1],
                "swap: buffers differ in kind/size");
    for (int k = 0; k < 2; k++) { void *t = x->p[k]; x->p[k] = y->p[k]; y->p[k] = t; }
    int b = x->borrowed; x->borrowed = y->borrowed; y->borrowed = b;     // ownership travels with the storage
    return 0;
}

int svl_scratch_node(svl_ctx *c, int k, svl_buf **out) {
    if (!c->psi_s[k]) SVL_TRY(svl_alloc(c, SVL_NODE_C, 0, 0, &c->psi_s[k]));
    *out = c->psi_s[k];
    return 0;
}

int svl_scratch_edge(svl_ctx *c, int k, svl_buf **out) {
    if (!c->ab_s[k]) SVL_TRY(svl_alloc(c, SVL_EDGE, 0, 0, &c->ab_s[k]));
    *out = c->ab_s[k];
    return 0;
}

// ----------------------------------------------------------------------------- seeded fields at scale (host)
// The reference seeds psi with numpy's LEGACY generator (np.random.seed + np.random.rand, svirl/vars/vars.py:93-109):
// MT19937 with 53-bit doubles (a >> 5, b >> 6).  Walking that stream in the interpreter costs ~15 ns per draw; at
// 32768^2 (2^31 draws) that dominated the build of a run.  Same recurrence here, on the state exported by
// numpy.random.RandomState(seed).get_state(), so the stream is bit-identical (tests/test_scale_host.py).
static inline void mt_twist(uint32_t *mt) {
    const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MA = 0x9908b0dfu;
    int kk = 0;
    for (; kk < 624 - 397; kk++) { uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MA : 0u); }
    for (; kk < 623; kk++) { uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MA : 0u); }
    uint32_t y = (mt[623] & UP) | (mt[0] & LO);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MA : 0u);
}
static inline uint32_t mt_next(uint32_t *mt, int *pos) {
    if (*pos >= 624) { mt_twist(mt); *pos = 0; }
    uint32_t y = mt[(*pos)++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}
extern "C" int svl_mt19937_doubles(uint32_t *key, int *pos, unsigned long long skip, double *out, unsigned long long n) {
    SVL_REQUIRE(key && pos && (out || n == 0) && *pos >= 0 && *pos <= 624, "bad MT19937 state");
    // skipping a double = skipping two 32-bit outputs; whole blocks of 624 need the twist only
    unsigned long long s = 2ull * skip;
    while (s > 0) {
        if (*pos >= 624) { mt_twist(key); *pos = 0; }
        unsigned long long take = (unsigned long long)(624 - *pos);
        if (take > s) take = s;
        *pos += (int)take; s -= take;
    }
    auto temper = [](uint32_t y) { y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18); return y; };
    const double inv53 = 1.0 / 9007199254740992.0;           // exact power of two: x * 2^-53 == x / 2^53 bit for bit
    unsigned long long k = 0;
    while (k < n) {
        if (*pos >= 624) { mt_twist(key); *pos = 0; }
        unsigned long long take = (unsigned long long)((624 - *pos) / 2);
        if (take > n - k) take = n - k;
        if (take == 0) {                                      // odd position: one double straddles two blocks
            const uint32_t a = mt_next(key, pos) >> 5, b = mt_next(key, pos) >> 6;
            out[k++] = (a * 67108864.0 + b) * inv53;
            continue;
        }
        const uint32_t *m = key + *pos;
        double *o = out + k;
        for (unsigned long long q = 0; q < take; q++) {       // branch-free: vectorises
            const uint32_t a = temper(m[2 * q]) >> 5, b = temper(m[2 * q + 1]) >> 6;
            o[q] = (a * 67108864.0 + b) * inv53;
        }
        *pos += (int)(2 * take); k += take;
    }
    return 0;
}

// psi0[n] = (1 - level*u1[n]) * exp(i*pi*level*(2*u2[n] - 1))   (svirl/vars/vars.py:106), evaluated like numpy does
// it for arrays -- the phase has a zero real part, so numpy's complex exp reduces to libm's cos / sin of
// (level*pi)*(2*u2 - 1), and the real-by-complex product to two roundings -- without numpy's six array temporaries.
// Bit-identical to the numpy expression on this platform (tests/test_scale_host.py); callers split n over threads.
extern "C" int svl_seeded_psi(const double *u1, const double *u2, unsigned long long n, double level, void *out,
                              int complex_bytes) {
    SVL_REQUIRE(u1 && u2 && out && (complex_bytes == 8 || complex_bytes == 16), "bad argument");
    const double lpi = level * 3.141592653589793;
    for (unsigned long long k = 0; k < n; k++) {
        const double m = 1.0 - level * u1[k];
        const double y = lpi * (2.0 * u2[k] - 1.0);
        double sn, cs;
        sincos(y, &sn, &cs);
        const double re = m * cs, im = m * sn;
        if (complex_bytes == 16) { ((double *)out)[2 * k] = re; ((double *)out)[2 * k + 1] = im; }
        else { ((float *)out)[2 * k] = (float)re; ((float *)out)[2 * k + 1] = (float)im; }
    }
    return 0;
}

// ----------------------------------------------------------------------------- self-test hook
// The library's own sincos (common.cuh) evaluated on n host values: lets the tests bound its
// error against the host libm without going through a solver.
template <typename R>
__global__ void k_debug_sincos(const double *x, double *s, double *c, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R ss, cc;
    sincos_r<R>((R)x[i], &ss, &cc);
    s[i] = (double)ss; c[i] = (double)cc;
}

extern "C" int svl_debug_sincos(svl_ctx *c, size_t n, const double *x, double *s_out, double *c_out) {
    SVL_REQUIRE(c && x && s_out && c_out && n > 0, "bad argument");
    double *d = nullptr;
    SVL_CHECK(cudaMalloc(&d, 3 * n * sizeof(double)));
    SVL_CHECK(cudaMemcpyAsync(d, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->rsize == 4) k_debug_sincos<float><<<svl_nblocks(n, 256), 256, 0, c->stream>>>(d, d + n, d + 2 * n, n);
    else k_debug_sincos<double><<<svl_nblocks(n, 256), 256, 0, c->stream>>>(d, d + n, d + 2 * n, n);
    SVL_CHECK(cudaGetLastError());
    SVL_CHECK(cudaMemcpyAsync(s_out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaMemcpyAsync(c_out, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    return 0;
}

// diagnostics: copy the per-launch trace records (4 words each) to the host; returns the count through n_out
extern "C" int svl_debug_trace(svl_ctx *c, unsigned long long *out, int max_records, int *n_out) {
    SVL_REQUIRE(c && out && n_out, "null argument");
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    int n = c->trace_n < max_records ? c->trace_n : max_records;
    if (n > 0) SVL_CHECK(cudaMemcpy(out, c->trace, (size_t)4 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    *n_out = n;
    return 0;
}
