// Tuning aid: one-channel linear warp with the 8 corners fetched by TWO texture-gather instructions (tld4) on a
// pitch-linear 2-D view of the volume (width = Z, height = X*Y) instead of 8 shared-memory loads from a TMA brick.
// tld4 returns the raw fp32 texels of the 2x2 footprint, so the arithmetic (weights, order) stays the library's.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -DDFM_EXACT_ORDER=0
//        -I multimodal-registration_b200/csrc -o scripts/exp/libtexprobe.so scripts/exp/tex_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dfm_common.cuh"

using namespace dfm;

__device__ __forceinline__ float ld_na(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int NX, int MODE>
__global__ void __launch_bounds__(256) k_warp_tex(const cudaTextureObject_t *__restrict__ texs, const float *__restrict__ field,
                                                  float *__restrict__ out, int X, int Y, int Z, int nxt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int z = (MODE & 4) ? blockIdx.x * 16 + (lane & 15) : blockIdx.x * 32 + lane;
    const int y = (MODE & 4) ? blockIdx.y * 16 + 2 * warp + (lane >> 4) : blockIdx.y * 8 + warp;
    const int b = blockIdx.z / nxt, x0 = (blockIdx.z - b * nxt) * NX;
    if (z >= Z || y >= Y) return;
    const cudaTextureObject_t tex = texs[b];
    const uint32_t N = (uint32_t)X * Y * Z, XS = (uint32_t)Y * Z;
    const float *fb = field + (size_t)b * 3 * N;
    float *ob = out + (size_t)b * N;
    const uint32_t vox0 = ((uint32_t)x0 * Y + y) * Z + z;
    const int mxi = X - 1, myi = Y - 1, mzi = Z - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    float l[3][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const uint32_t vox = vox0 + (uint32_t)min(i, X - 1 - x0) * XS;
        if (MODE & 1) { l[0][i] = ld_na(fb + vox); l[1][i] = ld_na(fb + N + vox); l[2][i] = ld_na(fb + 2 * (size_t)N + vox); }
        else { l[0][i] = __ldg(fb + vox); l[1][i] = __ldg(fb + N + vox); l[2][i] = __ldg(fb + 2 * (size_t)N + vox); }
    }
    float4 lo[NX], hi[NX];
    AxisF ax[NX], ay[NX], az[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        ax[i] = axis_fast(__fadd_rn((float)min(x0 + i, X - 1), l[0][i]), mxf, mxi);
        ay[i] = axis_fast(__fadd_rn((float)y, l[1][i]), myf, myi);
        az[i] = axis_fast(__fadd_rn((float)z, l[2][i]), mzf, mzi);
        const float u = (float)az[i].i1;                                 // texels i1-1, i1: footprint centre at i1
        const float v = (float)((ax[i].i1 - 1) * Y + ay[i].i1);
        lo[i] = tex2Dgather<float4>(tex, u, v, 0);
        hi[i] = tex2Dgather<float4>(tex, u, v + (float)Y, 0);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        if (x0 + i >= X) break;
        float w[8];
        tri_weights(ax[i], ay[i], az[i], w);
        // tld4 order: .w = (i, j), .z = (i+1, j), .x = (i, j+1), .y = (i+1, j+1); i along z, j along rows
        const float val[8] = {lo[i].w, lo[i].z, lo[i].x, lo[i].y, hi[i].w, hi[i].z, hi[i].x, hi[i].y};
        if (MODE & 2) __stcs(ob + vox0 + i * XS, tri_accumulate(w, val));
        else ob[vox0 + i * XS] = tri_accumulate(w, val);
    }
}

static cudaTextureObject_t *g_texs = nullptr;
static int g_ntex = 0;

extern "C" int texprobe_setup(const float *img, int B, int X, int Y, int Z) {
    cudaTextureObject_t h[256];
    if (B > 256) return -1;
    for (int b = 0; b < B; ++b) {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = (void *)(img + (size_t)b * X * Y * Z);
        rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        rd.res.pitch2D.width = Z;
        rd.res.pitch2D.height = (size_t)X * Y;
        rd.res.pitch2D.pitchInBytes = (size_t)Z * 4;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        cudaError_t e = cudaCreateTextureObject(&h[b], &rd, &td, nullptr);
        if (e != cudaSuccess) { fprintf(stderr, "texobj: %s\n", cudaGetErrorString(e)); return -2; }
    }
    if (!g_texs) cudaMalloc(&g_texs, 256 * sizeof(cudaTextureObject_t));
    cudaMemcpy(g_texs, h, B * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice);
    g_ntex = B;
    return 0;
}

template <int NXv, int MODE>
static void go(const float *field, float *out, int B, int X, int Y, int Z, int carve, cudaStream_t st) {
    const int nxt = (X + NXv - 1) / NXv;
    dim3 grid((MODE & 4) ? (Z + 15) / 16 : (Z + 31) / 32, (MODE & 4) ? (Y + 15) / 16 : (Y + 7) / 8, B * nxt);
    if (carve >= 0) cudaFuncSetAttribute(k_warp_tex<NXv, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    k_warp_tex<NXv, MODE><<<grid, 256, 0, st>>>(g_texs, field, out, X, Y, Z, nxt);
}

extern "C" int texprobe_warp(const float *field, float *out, int B, int X, int Y, int Z, int nx, int mode, int carve, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
#define GO(NXv)                                                                  \
    switch (mode) {                                                              \
        case 0: go<NXv, 0>(field, out, B, X, Y, Z, carve, st); break;            \
        case 1: go<NXv, 1>(field, out, B, X, Y, Z, carve, st); break;            \
        case 2: go<NXv, 2>(field, out, B, X, Y, Z, carve, st); break;            \
        case 3: go<NXv, 3>(field, out, B, X, Y, Z, carve, st); break;            \
        case 4: go<NXv, 4>(field, out, B, X, Y, Z, carve, st); break;            \
        case 7: go<NXv, 7>(field, out, B, X, Y, Z, carve, st); break;            \
        default: return -4;                                                      \
    }
    if (nx == 2) { GO(2) } else if (nx == 4) { GO(4) } else return -5;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "launch: %s\n", cudaGetErrorString(e)); return -3; }
    return 0;
}
