// Deterministic two-stage reductions.  Replaces svirl/parallel/reduction.py:40-173 and the
// kernels sum (svirl/cuda/block_reduction.h:81-92) / sum_v (svirl/cuda/reduction.h:57-82).
// Stage 1 (inside the producing kernel, or k_partial_* here) writes one partial per CTA;
// stage 2 is a single CTA that adds the partials in a fixed order, so results are
// run-to-run reproducible (no floating-point atomics anywhere).  Accumulation is in double
// for both precisions.
#include "common.cuh"

int svl_ensure_partials(svl_ctx *c, size_t n) {
    if (n <= c->partial_cap) return 0;
    if (c->partials) { SVL_CHECK(cudaStreamSynchronize(c->stream)); cudaFree(c->partials); c->partials = nullptr; }
    size_t cap = n + n / 4 + 1024;
    SVL_CHECK(cudaMalloc(&c->partials, cap * sizeof(double)));
    c->partial_cap = cap;
    return 0;
}

// out[k] = scale * sum_b partials[b*nv + k]; one CTA per component (fixed order inside a CTA: reproducible)
__global__ void __launch_bounds__(256) k_final_sum(const double *__restrict__ partials, int nblocks, int nv,
                                                   double scale, double *__restrict__ out) {
    __shared__ double sm[8];
    for (int k = blockIdx.x; k < nv; k += gridDim.x) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += 256) s += partials[(size_t)b * nv + k];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            s = threadIdx.x < 8 ? sm[threadIdx.x] : 0.0;
            s = warp_sum(s);
            if (threadIdx.x == 0) out[k] = s * scale;
        }
        __syncthreads();
    }
}

int svl_finish_sum(svl_ctx *c, int nblocks, int nv, double scale, double *out_host) {
    SVL_REQUIRE(nv <= 64, "too many reduction components");
    k_final_sum<<<nv, 256, 0, c->stream>>>(c->partials, nblocks, nv, scale, c->d_result);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    if (out_host) {
        SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SVL_CHECK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < nv; k++) out_host[k] = c->h_result[k];
    }
    return 0;
}

template <typename R>
__global__ void __launch_bounds__(256) k_partial_sum(const R *__restrict__ in, size_t n, double *partials) {
    double v[1] = {0.0};
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) v[0] += (double)in[i];
    block_sum_to_partials<1>(v, partials, blockIdx.x);
}

template <typename R>
__global__ void __launch_bounds__(256) k_partial_sum_v(const R *__restrict__ in, size_t nv, int ne, double *partials) {
    // vectors are stored one after another: in[j*ne + k] (svirl/cuda/reduction.h:68-71)
    for (int k = 0; k < ne; k++) {
        double v[1] = {0.0};
        for (size_t j = (size_t)blockIdx.x * 256 + threadIdx.x; j < nv; j += (size_t)gridDim.x * 256)
            v[0] += (double)in[j * ne + k];
        block_sum_to_partials<1>(v, partials + (size_t)k * gridDim.x, blockIdx.x);
        __syncthreads();
    }
}

extern "C" int svl_sum(svl_ctx *c, const svl_buf *in, size_t n, double *out) {
    SVL_REQUIRE(c && in && out, "null argument");
    SVL_REQUIRE(in->kind == SVL_FLAT && (in->esize == 4 || in->esize == 8), "svl_sum needs a real SVL_FLAT buffer");
    SVL_REQUIRE(n <= in->n, "n exceeds buffer");
    int nb = svl_nblocks(n ? n : 1, 256 * 8);
    if (nb > 1184) nb = 1184;   // 148 SMs x 8
    SVL_TRY(svl_ensure_partials(c, nb));
    if (in->esize == 4) k_partial_sum<float><<<nb, 256, 0, c->stream>>>((const float *)in->p[0], n, c->partials);
    else k_partial_sum<double><<<nb, 256, 0, c->stream>>>((const double *)in->p[0], n, c->partials);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    return svl_finish_sum(c, nb, 1, 1.0, out);
}

extern "C" int svl_sum_v(svl_ctx *c, const svl_buf *in, size_t nv, int ne, double *out) {
    SVL_REQUIRE(c && in && out, "null argument");
    SVL_REQUIRE(in->kind == SVL_FLAT && (in->esize == 4 || in->esize == 8), "svl_sum_v needs a real SVL_FLAT buffer");
    SVL_REQUIRE(ne >= 1 && ne <= 64 && nv * (size_t)ne <= in->n, "bad vector shape");
    int nb = svl_nblocks(nv ? nv : 1, 256 * 4);
    if (nb > 1184) nb = 1184;
    SVL_TRY(svl_ensure_partials(c, (size_t)nb * ne));
    if (in->esize == 4) k_partial_sum_v<float><<<nb, 256, 0, c->stream>>>((const float *)in->p[0], nv, ne, c->partials);
    else k_partial_sum_v<double><<<nb, 256, 0, c->stream>>>((const double *)in->p[0], nv, ne, c->partials);
    SVL_CHECK(cudaGetLastError());
    c->stat_launches += 1;
    // partials are laid out component-major here: [k][block]; reduce each component separately
    for (int k = 0; k < ne; k++) {
        k_final_sum<<<1, 256, 0, c->stream>>>(c->partials + (size_t)k * nb, nb, 1, 1.0, c->d_result + k);
        SVL_CHECK(cudaGetLastError());
    }
    SVL_CHECK(cudaMemcpyAsync(c->h_result, c->d_result, ne * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SVL_CHECK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < ne; k++) out[k] = c->h_result[k];
    return 0;
}
