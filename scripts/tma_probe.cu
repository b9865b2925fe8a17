// Stand-alone probe of 4-D TMA box loads at arbitrary coordinates (debug aid, not product code).
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, int nfloats,
                      float *out, int *status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *brick = (float *)smem_raw;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) brick[i] = -777.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nfloats * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(smem_u32(brick)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    uint32_t ok = 0, it = 0;
    for (; it < (1u << 22) && !ok; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    if (threadIdx.x == 0) { status[0] = ok; status[1] = it; }
    __syncthreads();
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = brick[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    if (argc < 13) { printf("usage: Z Y X C bz by bx bc c0 c1 c2 promo\n"); return 2; }
    int Z = atoi(argv[1]), Y = atoi(argv[2]), X = atoi(argv[3]), C = atoi(argv[4]);
    int bz = atoi(argv[5]), by = atoi(argv[6]), bx = atoi(argv[7]), bc = atoi(argv[8]);
    int c0 = atoi(argv[9]), c1 = atoi(argv[10]), c2 = atoi(argv[11]), promo = atoi(argv[12]);
    size_t n = (size_t)Z * Y * X * C;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *dout; int *dst;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int nf = bz * by * bx * bc;
    cudaMalloc(&dout, nf * 4); cudaMalloc(&dst, 8); cudaMemset(dst, 0, 8);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)C};
    cuuint64_t strides[3] = {(cuuint64_t)Z * 4, (cuuint64_t)Y * Z * 4, (cuuint64_t)X * Y * Z * 4};
    cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, (cuuint32_t)bc};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d ", (int)r);
    if (r) { printf("\n"); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, nf * 4);
    probe<<<1, 128, nf * 4>>>(tmap, c0, c1, c2, 0, nf, dout, dst);
    cudaError_t e = cudaDeviceSynchronize();
    int st[2] = {-1, -1};
    std::vector<float> o(nf, -1.f);
    if (e == cudaSuccess) { cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, nf * 4, cudaMemcpyDeviceToHost); }
    // verify
    long bad = 0;
    if (e == cudaSuccess)
        for (int cc = 0; cc < bc; ++cc) for (int x = 0; x < bx; ++x) for (int y = 0; y < by; ++y) for (int z = 0; z < bz; ++z) {
            int gz = c0 + z, gy = c1 + y, gx = c2 + x, gc = cc;
            float want = (gz >= 0 && gz < Z && gy >= 0 && gy < Y && gx >= 0 && gx < X && gc < C)
                             ? (float)((((size_t)gc * X + gx) * Y + gy) * Z + gz) : 0.f;
            float got = o[((cc * bx + x) * by + y) * bz + z];
            if (got != want) ++bad;
        }
    printf("sync=%s done=%d iters=%d mismatches=%ld\n", cudaGetErrorString(e), st[0], st[1], bad);
    return 0;
}
