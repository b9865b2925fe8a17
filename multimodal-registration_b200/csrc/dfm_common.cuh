// Shared device/host helpers of libdfm (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dfm.h"

namespace dfm {

// ---------------------------------------------------------------------------------------
// error plumbing (thread-local message, never throws)
// ---------------------------------------------------------------------------------------
char *err_buf();
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);

#define DFM_REQUIRE(cond, code, ...) \
    do {                             \
        if (!(cond)) return ::dfm::fail(code, __VA_ARGS__); \
    } while (0)

// planar out[b][c][n] = scale * in (3 channels; `in` planar or channels-last) -- dfm_layout.cu
// absmax (nullable): float per batch item, zeroed by the caller; receives max |out| of the item
int scale_copy_to_planar(const float *in, float *out, int B, size_t N, float scale, bool in_cl, float *absmax,
                         cudaStream_t st);

// running max |v|; a NaN counts as +inf (it sticks, and only disables the static halo)
__device__ __forceinline__ float absmax_fold(float m, float v) { return v != v ? __int_as_float(0x7f800000) : fmaxf(m, fabsf(v)); }
// block-wide max of a non-negative float, folded into *dst by one atomic per CTA (float bits order
// like ints for values >= 0).
// Must be reached by every thread of the block.
__device__ __forceinline__ void block_absmax_commit(float m, float *dst) {
    __shared__ int s_m;
    if (threadIdx.x == 0) s_m = 0;
    __syncthreads();
    atomicMax(&s_m, __float_as_int(m));
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int *>(dst), s_m);
}

int validate_grid(const char *who, int B, int X, int Y, int Z);

// TMA-brick path of out = scale*own + interp(scale*src, p + scale*own) -- dfm_brick.cu
bool brick_eligible(const float *src, const float *own, const float *out, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, unsigned flags);
// `bound` (nullable, device, one float per batch item) with `bscale`: bound[b] * bscale is an upper
// bound of |own| for item b; displacements below the brick's static halo skip the box reduction
int launch_ss_brick(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, float scale, int large_box, const float *bound, float bscale, cudaStream_t st);
// first SS step straight from a channels-last svf: optimistic static brick (dfm_brick.cu)
int launch_ss_first_cl(const float *svf, float *out, int B, int X, int Y, int Z, float scale, float *absmax,
                       cudaStream_t st);
// plane-marching SS step through a TMA ring of full-z rows (dfm_ss_march.cu); planar output.
// variant 0: halo 2, variant 1: halo 3.  first: v = scale*src, max|out| per item -> absmax (nullable).
// sel (nullable) + sel_mode 1/2: the CTAs of item b run iff (sel[b]*sel_scale < sel_thr) == (sel_mode == 1).
bool ss_march_eligible(const float *src, int X, int Y, int Z);
int launch_ss_march(const float *src, float *out, int B, int X, int Y, int Z, float scale, bool in_cl, bool first,
                    float *absmax, int variant, const float *sel, float sel_scale, float sel_thr, int sel_mode,
                    cudaStream_t st);
// channels-last multi-channel linear warp, lanes over channels (dfm_warp_cl.cu); DFM_EUNSUPPORTED if not applicable
int launch_warp_cl_fwd(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi, int X,
                       int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st);
int launch_warp_cl_bwd(const float *gout, const float *img, const float *field, float *gimg, float *gfield, int B,
                       int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags,
                       cudaStream_t st);
// channels-last warp backward with the Dice gradient formed on the fly: upstream g = coef[b][c][0] * y_true + coef[b][c][1]
int launch_warp_cl_dice_bwd(const float *y_true, const float *coef, const float *img, const float *field, float *gfield,
                            int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags,
                            cudaStream_t st);
// TMA-brick path of the one-channel linear image warp (returns DFM_EUNSUPPORTED if not applicable)
int launch_warp_brick(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X,
                      int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st);

// fused RescaleTransform + one-channel linear warp (returns DFM_EUNSUPPORTED if not applicable)
int launch_rescale_warp(const float *img, const float *half, float *out, const float *cx, const float *cy,
                        const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                        float pre, int has_fill, float fill, cudaStream_t st);

// the same with the image corners fetched by texture gathers and the field marched like the up-sampler -- dfm_warp_tex.cu
int launch_rescale_warp_tex(const float *img, const float *half, float *out, const float *cx, const float *cy,
                            const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                            float pre, int has_fill, float fill, cudaStream_t st);

// fused RescaleTransform + one-channel NEAREST warp of 4-byte elements on the same march (no textures) -- dfm_warp_tex.cu
int launch_rescale_warp_nearest(const void *img, const float *half, void *out, const float *cx, const float *cy,
                                const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                                float pre, int has_fill, uint32_t fill_bits, cudaStream_t st);
// stand-alone one-channel linear warp by texture gathers -- dfm_warp_tex.cu (DFM_EUNSUPPORTED if not applicable / not selected)
int launch_warp_tex(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X, int Y, int Z,
                    int has_fill, float fill, unsigned flags, cudaStream_t st);

// multi-channel planar linear warp through a TMA channel ring -- dfm_brick_mc.cu
int launch_warp_mc_fwd(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi, int X,
                       int Y, int Z, int has_fill, float fill, cudaStream_t st);
int launch_warp_mc_bwd_field(const float *gout, const float *img, const float *field, float *gfield, int B, int C,
                             int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, cudaStream_t st);

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// n / d for n*d < 2^32 via one umulhi (host checks the range)
struct FastDiv {
    uint32_t d, mul;
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.mul = d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, FastDiv f) {
    return f.d <= 1 ? n : __umulhi(n, f.mul);
}

// ---------------------------------------------------------------------------------------
// Arithmetic mode of the linear interpolation (compile-time; the Makefile builds both):
//   DFM_EXACT_ORDER=1 (libdfm_exact.so): every multiply and add of the reference's op chain
//       is a separately rounded fp32 operation in the reference's order -> results are
//       bit-identical to the oracle (TensorFlow's unfused Eigen element-wise kernels).
//   DFM_EXACT_ORDER=0 (libdfm.so, the default product): the same corner order, weights and
//       indices, but the 8-term accumulation uses fused multiply-adds, two voxels at a time on
//       Blackwell's packed FFMA2/FMUL2 pipe.  Differences are the removed intermediate
//       roundings (~1e-7 relative), far inside the 1e-5 rel / 1e-4 voxel parity bar.
// ---------------------------------------------------------------------------------------
#ifndef DFM_EXACT_ORDER
#define DFM_EXACT_ORDER 0
#endif

typedef unsigned long long u64_t;
__device__ __forceinline__ u64_t pk(float a, float b) {
    u64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk(u64_t r, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ u64_t mul2(u64_t a, u64_t b) {
    u64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64_t add2(u64_t a, u64_t b) {
    u64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64_t fma2(u64_t a, u64_t b, u64_t c) {
    u64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// ---------------------------------------------------------------------------------------
// one axis of neurite.utils.interpn's linear set-up, op for op:
//   loc0 = floor(loc); clipped = clip(loc); loc0c = clip(loc0); loc1 = clip(loc0c + 1)
//   w_lo = loc1 - clipped  (weight of the LOWER corner), w_hi = 1 - w_lo
// every fp32 op separately rounded (no FMA contraction anywhere in the sampling maths).
// ---------------------------------------------------------------------------------------
struct Axis {
    int i0, i1;
    float w0, w1;   // weights of corner i0 / i1
    float inb;      // 1 if 0 <= loc <= max (the clip passes gradient), else 0
};

__device__ __forceinline__ Axis axis_linear(float loc, float maxf) {
    Axis a;
    float fl = floorf(loc);
    float cl = fminf(fmaxf(loc, 0.f), maxf);
    float l0 = fminf(fmaxf(fl, 0.f), maxf);
    float l1 = fminf(__fadd_rn(l0, 1.f), maxf);
    a.i0 = (int)l0;
    a.i1 = (int)l1;
    a.w0 = __fsub_rn(l1, cl);
    a.w1 = __fsub_rn(1.f, a.w0);
    a.inb = (loc >= 0.f && loc <= maxf) ? 1.f : 0.f;
    return a;
}

// Exactness-preserving fast form of the same set-up (used by the shared-memory kernels):
//   clipped = clip(loc);  trunc(clipped) == clip(floor(loc)) for every finite loc, so
//   i1 = min(trunc(clipped) + 1, max) and w_lo = float(i1) - clipped are the reference's values.
//   The LOWER corner is addressed as i1 - 1; that differs from the reference's clip(floor(loc))
//   only when loc >= max, where w_lo is exactly 0 (0 * finite contributes nothing), which makes
//   the 8 corners one base address plus compile-time offsets.  Needs max >= 1.
struct AxisF {
    int i1;
    float w0, w1;
};
__device__ __forceinline__ AxisF axis_fast(float loc, float maxf, int maxi) {
    const float cl = fminf(fmaxf(loc, 0.f), maxf);
    AxisF a;
    a.i1 = min(__float2int_rz(cl) + 1, maxi);
    a.w0 = __fsub_rn((float)a.i1, cl);
    a.w1 = __fsub_rn(1.f, a.w0);
    return a;
}
__device__ __forceinline__ int axis_fast_i1(float loc, float maxf, int maxi) {
    return min(__float2int_rz(fminf(fmaxf(loc, 0.f), maxf)) + 1, maxi);
}
// the same set-up split in two, for kernels that clip once and keep the clipped location
__device__ __forceinline__ float axis_clip(float loc, float maxf) { return fminf(fmaxf(loc, 0.f), maxf); }
__device__ __forceinline__ int axis_clipped_i1(float cl, int maxi) { return min(__float2int_rz(cl) + 1, maxi); }
__device__ __forceinline__ AxisF axis_from_clipped(float cl, int maxi) {
    AxisF a;
    a.i1 = min(__float2int_rz(cl) + 1, maxi);
    a.w0 = __fsub_rn((float)a.i1, cl);
    a.w1 = __fsub_rn(1.f, a.w0);
    return a;
}
__device__ __forceinline__ void tri_weights(const AxisF &ax, const AxisF &ay, const AxisF &az, float (&w)[8]) {
    const float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
    const float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
    w[0] = __fmul_rn(w00, az.w0); w[1] = __fmul_rn(w00, az.w1);
    w[2] = __fmul_rn(w01, az.w0); w[3] = __fmul_rn(w01, az.w1);
    w[4] = __fmul_rn(w10, az.w0); w[5] = __fmul_rn(w10, az.w1);
    w[6] = __fmul_rn(w11, az.w0); w[7] = __fmul_rn(w11, az.w1);
}

// corner weights of two voxels at once (identical values in both modes: products only)
__device__ __forceinline__ void tri_weights_pair(const AxisF &ax, const AxisF &ay, const AxisF &az, const AxisF &bx,
                                                 const AxisF &by, const AxisF &bz, float (&wA)[8], float (&wB)[8]) {
#if DFM_EXACT_ORDER
    tri_weights(ax, ay, az, wA);
    tri_weights(bx, by, bz, wB);
#else
    const u64_t x0 = pk(ax.w0, bx.w0), x1 = pk(ax.w1, bx.w1), y0 = pk(ay.w0, by.w0), y1 = pk(ay.w1, by.w1);
    const u64_t z0 = pk(az.w0, bz.w0), z1 = pk(az.w1, bz.w1);
    const u64_t w00 = mul2(x0, y0), w01 = mul2(x0, y1), w10 = mul2(x1, y0), w11 = mul2(x1, y1);
    upk(mul2(w00, z0), wA[0], wB[0]); upk(mul2(w00, z1), wA[1], wB[1]);
    upk(mul2(w01, z0), wA[2], wB[2]); upk(mul2(w01, z1), wA[3], wB[3]);
    upk(mul2(w10, z0), wA[4], wB[4]); upk(mul2(w10, z1), wA[5], wB[5]);
    upk(mul2(w11, z0), wA[6], wB[6]); upk(mul2(w11, z1), wA[7], wB[7]);
#endif
}

// direct-gather variant: weights + flat offset of the lower corner; the 8 corners are then
// base + {0, 1, Z, Z+1, YZ, YZ+1, YZ+Z, YZ+Z+1}.  Needs X, Y, Z >= 2.
__device__ __forceinline__ uint32_t tri_setup_fast(float lx, float ly, float lz, int X, int Y, int Z, float (&w)[8]) {
    const AxisF ax = axis_fast(lx, (float)(X - 1), X - 1);
    const AxisF ay = axis_fast(ly, (float)(Y - 1), Y - 1);
    const AxisF az = axis_fast(lz, (float)(Z - 1), Z - 1);
    tri_weights(ax, ay, az, w);
    return ((uint32_t)(ax.i1 - 1) * (uint32_t)Y + (uint32_t)(ay.i1 - 1)) * (uint32_t)Z + (uint32_t)(az.i1 - 1);
}
// gather the 8 corners of one channel: p = channel base + lower-corner offset*es, es = element stride
__device__ __forceinline__ void gather8(const float *p, uint32_t gy, uint32_t gx, uint32_t es, float (&val)[8]) {
    const float *p01 = p + gy, *p10 = p + gx, *p11 = p10 + gy;
    val[0] = __ldg(p); val[1] = __ldg(p + es);
    val[2] = __ldg(p01); val[3] = __ldg(p01 + es);
    val[4] = __ldg(p10); val[5] = __ldg(p10 + es);
    val[6] = __ldg(p11); val[7] = __ldg(p11 + es);
}

// nearest: tf.round (half to even) on the unclipped location, then clip the integer
__device__ __forceinline__ int axis_nearest(float loc, int maxi) {
    float r = rintf(loc);
    // int32 cast of an out-of-range float is undefined in TF too; saturate for safety
    r = fminf(fmaxf(r, -2147483000.f), 2147483000.f);
    int i = (int)r;
    return min(max(i, 0), maxi);
}

// the 8 corner weights in itertools.product order, product taken left to right
struct Tri {
    size_t o[8];   // spatial offsets (x*YZ + y*Z + z), corner order 000,001,010,011,100,...
    float w[8];
};

__device__ __forceinline__ void tri_setup(float lx, float ly, float lz, int X, int Y, int Z,
                                          uint32_t (&off)[8], float (&w)[8]) {
    Axis ax = axis_linear(lx, (float)(X - 1));
    Axis ay = axis_linear(ly, (float)(Y - 1));
    Axis az = axis_linear(lz, (float)(Z - 1));
    const uint32_t YZ = (uint32_t)Y * (uint32_t)Z;
    uint32_t ox0 = (uint32_t)ax.i0 * YZ, ox1 = (uint32_t)ax.i1 * YZ;
    uint32_t oy0 = (uint32_t)ay.i0 * (uint32_t)Z, oy1 = (uint32_t)ay.i1 * (uint32_t)Z;
    uint32_t b00 = ox0 + oy0, b01 = ox0 + oy1, b10 = ox1 + oy0, b11 = ox1 + oy1;
    off[0] = b00 + az.i0; off[1] = b00 + az.i1;
    off[2] = b01 + az.i0; off[3] = b01 + az.i1;
    off[4] = b10 + az.i0; off[5] = b10 + az.i1;
    off[6] = b11 + az.i0; off[7] = b11 + az.i1;
    float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
    float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
    w[0] = __fmul_rn(w00, az.w0); w[1] = __fmul_rn(w00, az.w1);
    w[2] = __fmul_rn(w01, az.w0); w[3] = __fmul_rn(w01, az.w1);
    w[4] = __fmul_rn(w10, az.w0); w[5] = __fmul_rn(w10, az.w1);
    w[6] = __fmul_rn(w11, az.w0); w[7] = __fmul_rn(w11, az.w1);
}

// acc = ((0 + w0*v0) + w1*v1) + ...   in corner order; exact mode rounds every op separately
__device__ __forceinline__ float tri_accumulate(const float (&w)[8], const float (&v)[8]) {
    float acc = __fmul_rn(w[0], v[0]);
#if DFM_EXACT_ORDER
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fadd_rn(acc, __fmul_rn(w[k], v[k]));
#else
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fmaf_rn(w[k], v[k], acc);
#endif
    return acc;
}

// the same for two voxels A and B at once (packed pipe in fast mode)
__device__ __forceinline__ void tri_accumulate_pair(const float (&wA)[8], const float (&wB)[8], const float (&vA)[8],
                                                    const float (&vB)[8], float &a, float &b) {
#if DFM_EXACT_ORDER
    a = tri_accumulate(wA, vA);
    b = tri_accumulate(wB, vB);
#else
    u64_t acc = mul2(pk(wA[0], wB[0]), pk(vA[0], vB[0]));
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = fma2(pk(wA[k], wB[k]), pk(vA[k], vB[k]), acc);
    upk(acc, a, b);
#endif
}

__device__ __forceinline__ bool oob3(float lx, float ly, float lz, int X, int Y, int Z) {
    return lx < 0.f || lx > (float)(X - 1) || ly < 0.f || ly > (float)(Y - 1) || lz < 0.f ||
           lz > (float)(Z - 1);
}

}  // namespace dfm
