// TMA-staged shared-memory bricks for the gather-heavy self-warp of a scaling-and-squaring
// step (and compose):   out = scale*own + interp(scale*src, p + scale*own),  planar fp32.
//
// A CTA owns a tile of TX x TY x TZ = 8 x 8 x 32 output voxels (warp = one y row, lane = z,
// each thread walks the 8 x planes).  Pass 1 loads the tile's own vectors (coalesced), forms
// the sample locations and block-reduces their integer bounding box.  One elected thread then
// issues a single 4-D TMA box load {BZ, BY, BX, 3 components} whose ORIGIN is that bounding
// box's corner -- so the brick follows the displacement, and its size only has to cover the
// tile plus the local deformation, not the displacement magnitude.  Pass 2 gathers the 24
// corner values per voxel from shared memory (32-bit addressing, one wavefront per request
// instead of two L1 lines) with the reference's op order.  A thread whose corners fall outside
// the brick (strong local deformation) falls back to global gathers, so the result never
// depends on the brick size.  TMA zero-fills out-of-volume elements; they are never read
// because corner indices are clamped to the volume first (edge clamp of the reference).
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "dfm_common.cuh"

namespace dfm {

// ------------------------------- PTX helpers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    // bounded spin: a barrier that never completes traps instead of hanging the GPU
    for (uint32_t it = 0; !mbar_try_wait(bar, phase); ++it)
        if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

constexpr int TX = 8, TY = 8, TZ = 32;

template <int BX, int BY, int BZ>
__global__ void __launch_bounds__(256)
k_ss_brick(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src,
           const float *__restrict__ own, float *__restrict__ out, int Xs, int Ys, int Zs, int X, int Y,
           int Z, float scale, int nzt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *brick = reinterpret_cast<float *>(smem_raw);   // [3][BX][BY][BZ]
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_min[3], s_max[3];
    constexpr int CS = BX * BY * BZ;                      // component stride in the brick

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int z = zt * TZ + lane, y = yt * TY + warp, x0 = blockIdx.y * TX;
    const bool ok_yz = (z < Z) && (y < Y);
    const uint32_t N = (uint32_t)X * Y * Z, Ns = (uint32_t)Xs * Ys * Zs;
    const float *ownb = own + (size_t)blockIdx.z * 3 * N;
    const float *srcb = src + (size_t)blockIdx.z * 3 * Ns;
    float *outb = out + (size_t)blockIdx.z * 3 * N;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        s_min[0] = s_min[1] = s_min[2] = INT_MAX;
        s_max[0] = s_max[1] = s_max[2] = INT_MIN;
    }
    __syncthreads();

    // ---- pass 1: own vectors, sample locations, bounding box of the corner indices --------
    const float mxf = (float)(Xs - 1), myf = (float)(Ys - 1), mzf = (float)(Zs - 1);
    const float fy = (float)y, fz = (float)z;
    float v[3][TX];
    int mn0 = INT_MAX, mn1 = INT_MAX, mn2 = INT_MAX, mx0 = INT_MIN, mx1 = INT_MIN, mx2 = INT_MIN;
#pragma unroll
    for (int i = 0; i < TX; ++i) {
        const int x = x0 + i;
        if (ok_yz && x < X) {
            const uint32_t vox = ((uint32_t)x * Y + y) * Z + z;
            v[0][i] = __fmul_rn(scale, __ldg(ownb + vox));
            v[1][i] = __fmul_rn(scale, __ldg(ownb + N + vox));
            v[2][i] = __fmul_rn(scale, __ldg(ownb + 2 * (size_t)N + vox));
        } else {
            v[0][i] = v[1][i] = v[2][i] = 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < TX; ++i) {
        const int x = x0 + i;
        if (ok_yz && x < X) {
            const float lx = __fadd_rn((float)x, v[0][i]), ly = __fadd_rn(fy, v[1][i]), lz = __fadd_rn(fz, v[2][i]);
            const int ix = (int)fminf(fmaxf(floorf(lx), 0.f), mxf);
            const int iy = (int)fminf(fmaxf(floorf(ly), 0.f), myf);
            const int iz = (int)fminf(fmaxf(floorf(lz), 0.f), mzf);
            mn0 = min(mn0, ix); mn1 = min(mn1, iy); mn2 = min(mn2, iz);
            mx0 = max(mx0, ix); mx1 = max(mx1, iy); mx2 = max(mx2, iz);
        }
    }
    mn0 = __reduce_min_sync(0xffffffffu, mn0); mn1 = __reduce_min_sync(0xffffffffu, mn1);
    mn2 = __reduce_min_sync(0xffffffffu, mn2); mx0 = __reduce_max_sync(0xffffffffu, mx0);
    mx1 = __reduce_max_sync(0xffffffffu, mx1); mx2 = __reduce_max_sync(0xffffffffu, mx2);
    if (lane == 0 && mn0 != INT_MAX) {
        atomicMin(&s_min[0], mn0); atomicMin(&s_min[1], mn1); atomicMin(&s_min[2], mn2);
        atomicMax(&s_max[0], mx0); atomicMax(&s_max[1], mx1); atomicMax(&s_max[2], mx2);
    }
    __syncthreads();
    // measured on sm_100a: the innermost TMA coordinate must be 16-byte aligned (a multiple of 4
    // floats) or the copy raises an illegal-instruction fault -> align the box origin down in z
    const int ox = s_min[0], oy = s_min[1], oz = s_min[2] & ~3;
    if (ox == INT_MAX) return;                        // tile entirely outside the volume (uniform)
    // the upper corner index is min(i0 + 1, max): the brick must reach one past the largest i0
    const bool fits = (min(s_max[0] + 1, Xs - 1) - ox < BX) && (min(s_max[1] + 1, Ys - 1) - oy < BY) &&
                      (min(s_max[2] + 1, Zs - 1) - oz < BZ);

    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 3u * CS * sizeof(float));
        tma_load_4d(brick, &tmap, &bar, oz, oy, ox, (int)blockIdx.z * 3);
    }
    mbar_wait(&bar, 0);

    // ---- pass 2: gather from the brick, reference op order --------------------------------
#pragma unroll
    for (int i = 0; i < TX; ++i) {
        const int x = x0 + i;
        if (!(ok_yz && x < X)) continue;
        const float v0 = v[0][i], v1 = v[1][i], v2 = v[2][i];
        const Axis ax = axis_linear(__fadd_rn((float)x, v0), mxf);
        const Axis ay = axis_linear(__fadd_rn(fy, v1), myf);
        const Axis az = axis_linear(__fadd_rn(fz, v2), mzf);
        float w[8];
        {
            const float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
            const float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
            w[0] = __fmul_rn(w00, az.w0); w[1] = __fmul_rn(w00, az.w1);
            w[2] = __fmul_rn(w01, az.w0); w[3] = __fmul_rn(w01, az.w1);
            w[4] = __fmul_rn(w10, az.w0); w[5] = __fmul_rn(w10, az.w1);
            w[6] = __fmul_rn(w11, az.w0); w[7] = __fmul_rn(w11, az.w1);
        }
        float a0, a1, a2;
        const bool inside = fits || ((ax.i1 - ox < BX) && (ay.i1 - oy < BY) && (az.i1 - oz < BZ));
        if (inside) {
            const int bx0 = (ax.i0 - ox) * (BY * BZ), bx1 = (ax.i1 - ox) * (BY * BZ);
            const int by0 = (ay.i0 - oy) * BZ, by1 = (ay.i1 - oy) * BZ;
            const int bz0 = az.i0 - oz, bz1 = az.i1 - oz;
            const float *q00 = brick + bx0 + by0, *q01 = brick + bx0 + by1;
            const float *q10 = brick + bx1 + by0, *q11 = brick + bx1 + by1;
            float val[8];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                val[0] = q00[c * CS + bz0]; val[1] = q00[c * CS + bz1];
                val[2] = q01[c * CS + bz0]; val[3] = q01[c * CS + bz1];
                val[4] = q10[c * CS + bz0]; val[5] = q10[c * CS + bz1];
                val[6] = q11[c * CS + bz0]; val[7] = q11[c * CS + bz1];
                const float a = tri_accumulate(w, val);
                if (c == 0) a0 = a; else if (c == 1) a1 = a; else a2 = a;
            }
        } else {
            const uint32_t YZ = (uint32_t)Ys * Zs;
            uint32_t off[8];
            const uint32_t b00 = ax.i0 * YZ + ay.i0 * Zs, b01 = ax.i0 * YZ + ay.i1 * Zs;
            const uint32_t b10 = ax.i1 * YZ + ay.i0 * Zs, b11 = ax.i1 * YZ + ay.i1 * Zs;
            off[0] = b00 + az.i0; off[1] = b00 + az.i1; off[2] = b01 + az.i0; off[3] = b01 + az.i1;
            off[4] = b10 + az.i0; off[5] = b10 + az.i1; off[6] = b11 + az.i0; off[7] = b11 + az.i1;
            float val[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = __ldg(srcb + off[k]);
            a0 = tri_accumulate(w, val);
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = __ldg(srcb + Ns + off[k]);
            a1 = tri_accumulate(w, val);
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] = __ldg(srcb + 2 * (size_t)Ns + off[k]);
            a2 = tri_accumulate(w, val);
        }
        const uint32_t vox = ((uint32_t)x * Y + y) * Z + z;
        outb[vox] = __fadd_rn(v0, __fmul_rn(scale, a0));
        outb[N + vox] = __fadd_rn(v1, __fmul_rn(scale, a1));
        outb[2 * (size_t)N + vox] = __fadd_rn(v2, __fmul_rn(scale, a2));
    }
}

// ------------------------------- host side -----------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

bool brick_eligible(const float *src, const float *own, const float *out, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, unsigned flags) {
    if (flags & (DFM_FIELD_IN_CL | DFM_FIELD_OUT_CL)) return false;
    static const bool disabled = getenv("DFM_NO_BRICK") != nullptr;
    if (disabled) return false;
    if (Zs % 4 != 0 || !aligned16(src)) return false;        // TMA: 16-byte global strides
    if (Xs < 2 || Ys < 2 || Zs < 4) return false;
    (void)own; (void)out; (void)X; (void)Y; (void)Z;
    return encode_fn() != nullptr;
}

template <int BX, int BY, int BZ>
static int launch_brick_t(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                          int Y, int Z, float scale, cudaStream_t st) {
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {(cuuint64_t)Zs, (cuuint64_t)Ys, (cuuint64_t)Xs, (cuuint64_t)B * 3};
    const cuuint64_t strides[3] = {(cuuint64_t)Zs * 4, (cuuint64_t)Ys * Zs * 4, (cuuint64_t)Xs * Ys * Zs * 4};
    const cuuint32_t box[4] = {BZ, BY, BX, 3};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)src, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return DFM_EUNSUPPORTED;          // caller falls back to direct gathers
    constexpr size_t smem = 3ull * BX * BY * BZ * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_ss_brick<BX, BY, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_ss_brick smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int nzt = (Z + TZ - 1) / TZ, nyt = (Y + TY - 1) / TY, nxt = (X + TX - 1) / TX;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    if (getenv("DFM_BRICK_DEBUG")) fprintf(stderr, "dfm: k_ss_brick<%d,%d,%d> grid (%d,%d,%d) smem %zu scale %g\n", BX, BY, BZ, grid.x, grid.y, grid.z, smem, scale);
    k_ss_brick<BX, BY, BZ><<<grid, block, smem, st>>>(tmap, src, own, out, Xs, Ys, Zs, X, Y, Z, scale, nzt);
    return check_launch("k_ss_brick");
}

int launch_ss_brick(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, float scale, int large_box, cudaStream_t st) {
    // z extent = 32 (tile) + 1 (upper corner) + 3 (origin alignment) + deformation slack
    if (large_box) return launch_brick_t<12, 12, 52>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, st);
    return launch_brick_t<10, 10, 40>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, st);
}

}  // namespace dfm
