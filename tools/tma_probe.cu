// Bring-up probe for 2-D TMA box loads (not part of the product): one case per process.
//   tma_probe <elem: f32|f64|u8> <width> <rows> <pitch_elems> <box_w> <box_h> <c0> <c1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, uint32_t bytes, unsigned char *out, int *status) {
    extern __shared__ unsigned char raw[];
    unsigned char *sm = (unsigned char *)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    uint64_t *bar = (uint64_t *)(sm + 65536);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(d), "l"((uint64_t)&tm), "r"(c0), "r"(c1), "r"(b) : "memory");
    }
    long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(b) : "memory");
        if (clock64() - t0 > 2000000000ll) break;
    }
    if (threadIdx.x == 0) *status = ok ? 1 : -1;
    __syncthreads();
    if (ok) for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char **argv) {
    if (argc < 9) return 2;
    const char *el = argv[1];
    size_t width = atol(argv[2]), rows = atol(argv[3]), pitch = atol(argv[4]);
    int bw = atoi(argv[5]), bh = atoi(argv[6]), c0 = atoi(argv[7]), c1 = atoi(argv[8]);
    int es = !strcmp(el, "f32") ? 4 : !strcmp(el, "f64") ? 8 : 1;
    CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
    unsigned char *g; size_t gb = rows * pitch * es;
    cudaMalloc(&g, gb);
    unsigned char *h = (unsigned char *)malloc(gb);
    for (size_t i = 0; i < gb; i++) h[i] = (unsigned char)(1 + (i * 7) % 250);
    cudaMemcpy(g, h, gb, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[2] = {width, rows}; cuuint64_t strides[1] = {pitch * es};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, est[2] = {1, 1};
    CUresult r = ((PFN_encodeTiled)p)(&tm, dt, 2, g, dims, strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("case %s w=%zu rows=%zu pitch=%zu box=%dx%d at (%d,%d): encode=%d ", el, width, rows, pitch, bw, bh, c0, c1, (int)r);
    if (r) { printf("\n"); return 0; }
    uint32_t bytes = (uint32_t)bw * bh * es;
    unsigned char *out; int *st;
    cudaMalloc(&out, bytes); cudaMemset(out, 0xEE, bytes); cudaMalloc(&st, 4); cudaMemset(st, 0, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 256);
    k<<<1, 128, 65536 + 256>>>(tm, c0, c1, bytes, out, st);
    cudaError_t e = cudaDeviceSynchronize();
    int sth = 0;
    if (e == cudaSuccess) cudaMemcpy(&sth, st, 4, cudaMemcpyDeviceToHost);
    printf("run=%s status=%d ", cudaGetErrorString(e), sth);
    if (e == cudaSuccess && sth == 1) {
        unsigned char *ho = (unsigned char *)malloc(bytes);
        cudaMemcpy(ho, out, bytes, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int rr = 0; rr < bh; rr++)
            for (int cc = 0; cc < bw * es; cc++) {
                long gx = (long)c0 * es + cc, gy = c1 + rr;
                unsigned char want = 0;
                if (gx >= 0 && gx < (long)(width * es) && gy >= 0 && gy < (long)rows) want = h[gy * pitch * es + gx];
                if (ho[(size_t)rr * bw * es + cc] != want) bad++;
            }
        printf("mismatches=%zu", bad);
    }
    printf("\n");
    return 0;
}
