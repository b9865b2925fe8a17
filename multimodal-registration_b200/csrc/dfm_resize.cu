// ne.utils.resize / rescale_dense_transform: resample C channels onto a separable coordinate
// grid given by per-axis tables (tf.linspace(0, n_in-1, n_out) in the reference).
//   out[b,c,jx,jy,jz] = post * interp(pre * in[b,c], (cx[jx], cy[jy], cz[jz]))
// The x/y axis set-up is uniform along an output row, so a thread computes it once for its
// VEC consecutive z outputs; the input (1/8 of the output for a x2 upsample) stays in L1/L2.
#include "dfm_common.cuh"

namespace dfm {

template <int VEC, int INTERP, bool IN_CL, bool OUT_CL>
__global__ void __launch_bounds__(256)
k_resize(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ cx,
         const float *__restrict__ cy, const float *__restrict__ cz, int C, int Xi, int Yi, int Zi,
         int Xo, int Yo, int Zo, float pre, float post, FastDiv zvdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t jy = fast_div(p, zvdiv);
    const uint32_t jz = (p - jy * zvdiv.d) * VEC;
    const uint32_t jx = blockIdx.y;
    const size_t No = (size_t)Xo * Yo * Zo, Ni = (size_t)Xi * Yi * Zi;
    const size_t vox = ((size_t)jx * Yo + jy) * Zo + jz;
    const float *ib = in + (size_t)blockIdx.z * C * Ni;
    float *ob = out + (size_t)blockIdx.z * C * No;
    const float lx = __ldg(cx + jx), ly = __ldg(cy + jy);

    if (INTERP == DFM_LINEAR) {
        const Axis ax = axis_linear(lx, (float)(Xi - 1));
        const Axis ay = axis_linear(ly, (float)(Yi - 1));
        const uint32_t YZ = (uint32_t)Yi * Zi;
        const uint32_t b00 = ax.i0 * YZ + ay.i0 * Zi, b01 = ax.i0 * YZ + ay.i1 * Zi;
        const uint32_t b10 = ax.i1 * YZ + ay.i0 * Zi, b11 = ax.i1 * YZ + ay.i1 * Zi;
        const float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
        const float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
        uint32_t off[VEC][8];
        float w[VEC][8];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const Axis az = axis_linear(__ldg(cz + jz + i), (float)(Zi - 1));
            off[i][0] = b00 + az.i0; off[i][1] = b00 + az.i1; off[i][2] = b01 + az.i0; off[i][3] = b01 + az.i1;
            off[i][4] = b10 + az.i0; off[i][5] = b10 + az.i1; off[i][6] = b11 + az.i0; off[i][7] = b11 + az.i1;
            w[i][0] = __fmul_rn(w00, az.w0); w[i][1] = __fmul_rn(w00, az.w1);
            w[i][2] = __fmul_rn(w01, az.w0); w[i][3] = __fmul_rn(w01, az.w1);
            w[i][4] = __fmul_rn(w10, az.w0); w[i][5] = __fmul_rn(w10, az.w1);
            w[i][6] = __fmul_rn(w11, az.w0); w[i][7] = __fmul_rn(w11, az.w1);
        }
        for (int c = 0; c < C; ++c) {
            float r[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float val[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float t = IN_CL ? __ldg(ib + (size_t)off[i][k] * C + c) : __ldg(ib + (size_t)c * Ni + off[i][k]);
                    val[k] = __fmul_rn(pre, t);
                }
                r[i] = __fmul_rn(post, tri_accumulate(w[i], val));
            }
            if (OUT_CL) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) ob[(vox + i) * C + c] = r[i];
            } else if (VEC == 4) {
                *reinterpret_cast<float4 *>(ob + (size_t)c * No + vox) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
                ob[(size_t)c * No + vox] = r[0];
            }
        }
    } else {
        const uint32_t bxy = ((uint32_t)axis_nearest(lx, Xi - 1) * Yi + axis_nearest(ly, Yi - 1)) * Zi;
        uint32_t off[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) off[i] = bxy + axis_nearest(__ldg(cz + jz + i), Zi - 1);
        for (int c = 0; c < C; ++c) {
            float r[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float t = IN_CL ? __ldg(ib + (size_t)off[i] * C + c) : __ldg(ib + (size_t)c * Ni + off[i]);
                r[i] = __fmul_rn(post, __fmul_rn(pre, t));
            }
            if (OUT_CL) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) ob[(vox + i) * C + c] = r[i];
            } else if (VEC == 4) {
                *reinterpret_cast<float4 *>(ob + (size_t)c * No + vox) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
                ob[(size_t)c * No + vox] = r[0];
            }
        }
    }
}

template <int VEC, int INTERP>
static int launch_resize(const float *in, float *out, const float *cx, const float *cy, const float *cz, int B,
                         int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre, float post,
                         unsigned flags, cudaStream_t st) {
    const uint32_t zv = Zo / VEC, plane = (uint32_t)Yo * zv;
    dim3 grid((plane + 255) / 256, Xo, B), block(256);
    FastDiv fd = make_fastdiv(zv);
    const bool icl = flags & DFM_FIELD_IN_CL, ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(I, O) k_resize<VEC, INTERP, I, O><<<grid, block, 0, st>>>( \
        in, out, cx, cy, cz, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, fd, plane)
    if (icl) { if (ocl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (ocl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_resize_fwd");
}

// ---------------------------------------------------------------------------------------
// adjoint (gather form): gin[i] = pre*post * sum_{jx in Rx(ix)} sum_{jy} sum_{jz} wx wy wz gout[j]
// where wa(j, i) = weight that output j puts on input index i along axis a.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float axis_weight_on(float loc, float maxf, int i) {
    const Axis a = axis_linear(loc, maxf);
    float w = 0.f;
    if (a.i0 == i) w += a.w0;
    if (a.i1 == i) w += a.w1;
    return w;
}

__global__ void __launch_bounds__(128)
k_resize_bwd(const float *__restrict__ gout, float *__restrict__ gin, const float *__restrict__ cx,
             const float *__restrict__ cy, const float *__restrict__ cz, const int *__restrict__ xlo,
             const int *__restrict__ xhi, const int *__restrict__ ylo, const int *__restrict__ yhi,
             const int *__restrict__ zlo, const int *__restrict__ zhi, int C, int Xi, int Yi, int Zi,
             int Xo, int Yo, int Zo, float s, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t iy = fast_div(p, zdiv);
    const uint32_t iz = p - iy * zdiv.d;
    const uint32_t ix = blockIdx.y;
    const uint32_t bc = blockIdx.z;   // b*C + c
    const size_t No = (size_t)Xo * Yo * Zo, Ni = (size_t)Xi * Yi * Zi;
    const float *gb = gout + (size_t)bc * No;
    const int x0 = xlo[ix], x1 = xhi[ix], y0 = ylo[iy], y1 = yhi[iy], z0 = zlo[iz], z1 = zhi[iz];
    float acc = 0.f;
    for (int jx = x0; jx < x1; ++jx) {
        const float wx = axis_weight_on(__ldg(cx + jx), (float)(Xi - 1), ix);
        for (int jy = y0; jy < y1; ++jy) {
            const float wxy = wx * axis_weight_on(__ldg(cy + jy), (float)(Yi - 1), iy);
            const float *row = gb + ((size_t)jx * Yo + jy) * Zo;
            float accz = 0.f;
            for (int jz = z0; jz < z1; ++jz)
                accz = fmaf(axis_weight_on(__ldg(cz + jz), (float)(Zi - 1), iz), __ldg(row + jz), accz);
            acc = fmaf(wxy, accz, acc);
        }
    }
    gin[(size_t)bc * Ni + ((size_t)ix * Yi + iy) * Zi + iz] = s * acc;
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_resize_fwd(const float *in, float *out, const float *cx, const float *cy, const float *cz,
                              int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre,
                              float post, int interp, unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xo >= 0 && Yo >= 0 && Zo >= 0, DFM_EINVAL,
                "dfm_resize_fwd: bad shape B=%d C=%d in=(%d,%d,%d) out=(%d,%d,%d)", B, C, Xi, Yi, Zi, Xo, Yo, Zo);
    DFM_REQUIRE(B <= 65535 && Xo <= 65535, DFM_EINVAL, "dfm_resize_fwd: B and Xo must be <= 65535");
    DFM_REQUIRE((uint64_t)Xo * Yo * Zo < (1ull << 31) && (uint64_t)Xi * Yi * Zi < (1ull << 31), DFM_EINVAL,
                "dfm_resize_fwd: volume too large (>= 2^31 voxels)");
    DFM_REQUIRE((uint64_t)Yo * Zo * (uint64_t)Zo < (1ull << 32), DFM_EINVAL, "dfm_resize_fwd: Yo*Zo*Zo must be < 2^32");
    DFM_REQUIRE(interp == DFM_LINEAR || interp == DFM_NEAREST, DFM_EINVAL, "dfm_resize_fwd: interp %d", interp);
    if (B == 0 || Xo == 0 || Yo == 0 || Zo == 0) return DFM_OK;
    DFM_REQUIRE(in && out && cx && cy && cz, DFM_EINVAL, "dfm_resize_fwd: null pointer");
    DFM_REQUIRE(in != out, DFM_EINVAL, "dfm_resize_fwd: out must not alias in");
    if (C == 1) flags &= ~(DFM_FIELD_IN_CL | DFM_FIELD_OUT_CL);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (Zo % 4 == 0) && aligned16(out);
    if (interp == DFM_LINEAR)
        return vec4 ? launch_resize<4, DFM_LINEAR>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st)
                    : launch_resize<1, DFM_LINEAR>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st);
    return vec4 ? launch_resize<4, DFM_NEAREST>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st)
                : launch_resize<1, DFM_NEAREST>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st);
}

extern "C" int dfm_resize_bwd(const float *gout, float *gin, const float *cx, const float *cy, const float *cz,
                              const int *xlo, const int *xhi, const int *ylo, const int *yhi, const int *zlo,
                              const int *zhi, int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo,
                              float pre, float post, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xo >= 0 && Yo >= 0 && Zo >= 0, DFM_EINVAL,
                "dfm_resize_bwd: bad shape");
    DFM_REQUIRE((uint64_t)B * C <= 65535 && Xi <= 65535, DFM_EINVAL, "dfm_resize_bwd: B*C and Xi must be <= 65535");
    DFM_REQUIRE((uint64_t)Xo * Yo * Zo < (1ull << 31) && (uint64_t)Xi * Yi * Zi < (1ull << 31), DFM_EINVAL,
                "dfm_resize_bwd: volume too large");
    DFM_REQUIRE((uint64_t)Yi * Zi * (uint64_t)Zi < (1ull << 32), DFM_EINVAL, "dfm_resize_bwd: Yi*Zi*Zi must be < 2^32");
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(gout && gin && cx && cy && cz && xlo && xhi && ylo && yhi && zlo && zhi, DFM_EINVAL,
                "dfm_resize_bwd: null pointer");
    const uint32_t plane = (uint32_t)Yi * Zi;
    dim3 grid((plane + 127) / 128, Xi, B * C), block(128);
    k_resize_bwd<<<grid, block, 0, (cudaStream_t)stream>>>(gout, gin, cx, cy, cz, xlo, xhi, ylo, yhi, zlo, zhi, C,
                                                          Xi, Yi, Zi, Xo, Yo, Zo, pre * post, make_fastdiv(Zi), plane);
    return check_launch("dfm_resize_bwd");
}
