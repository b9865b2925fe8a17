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
int scale_copy_to_planar(const float *in, float *out, int B, size_t N, float scale, bool in_cl,
                         cudaStream_t st);

int validate_grid(const char *who, int B, int X, int Y, int Z);

// TMA-brick path of out = scale*own + interp(scale*src, p + scale*own) -- dfm_brick.cu
bool brick_eligible(const float *src, const float *own, const float *out, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, unsigned flags);
int launch_ss_brick(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, float scale, int large_box, cudaStream_t st);

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

// acc = ((0 + w0*v0) + w1*v1) + ...   in corner order, each op rounded separately
__device__ __forceinline__ float tri_accumulate(const float (&w)[8], const float (&v)[8]) {
    float acc = __fmul_rn(w[0], v[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fadd_rn(acc, __fmul_rn(w[k], v[k]));
    return acc;
}

__device__ __forceinline__ bool oob3(float lx, float ly, float lz, int X, int Y, int Z) {
    return lx < 0.f || lx > (float)(X - 1) || ly < 0.f || ly > (float)(Y - 1) || lz < 0.f ||
           lz > (float)(Z - 1);
}

}  // namespace dfm
