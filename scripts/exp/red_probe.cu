// Microbenchmark (tuning aid): throughput of coalesced global float reductions, scalar vs .v2 vs .v4
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/exp/red_probe.cu -o scripts/exp/red_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void k_red(float *dst, size_t nvec, int reps, int shift) {
    // element e of the grid-stride range adds to dst[(e + r*shift) * V .. +V): consecutive lanes -> consecutive vectors
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nvec; i += stride) {
        for (int r = 0; r < reps; ++r) {
            size_t j = (i + (size_t)r * shift) % nvec;
            float *p = dst + j * V;
            if (V == 1) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
            if (V == 2) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(1.0f) : "memory");
            if (V == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
        }
    }
}

template <int V>
void run(float *d, size_t nfloat, int reps, int shift) {
    size_t nvec = nfloat / V;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) k_red<V><<<148 * 8, 256>>>(d, nvec, reps, shift);
    cudaEventRecord(a);
    for (int w = 0; w < 5; ++w) k_red<V><<<148 * 8, 256>>>(d, nvec, reps, shift);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("V=%d nfloat=%zu reps=%d shift=%d: %.3f ms  %.1f Gfloat-adds/s  %.1f Gops/s  [%s]\n", V, nfloat, reps, shift, ms,
           (double)nvec * V * reps / ms / 1e6, (double)nvec * reps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t n = 3ull * 80 * 80 * 96 * 2 * 4;      // a B=2 SS gradient field, x4 for the padded layout
    float *d; cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    for (int shift : {1, 97, 7681}) {
        run<1>(d, n, 8, shift);
        run<2>(d, n, 8, shift);
        run<4>(d, n, 8, shift);
    }
    // larger working set (beyond L2)
    const size_t big = 1ull << 28;
    float *e; cudaMalloc(&e, big * 4); cudaMemset(e, 0, big * 4);
    run<1>(e, big, 4, 97); run<2>(e, big, 4, 97); run<4>(e, big, 4, 97);
    return 0;
}
