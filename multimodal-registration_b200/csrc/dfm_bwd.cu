// Backward passes of the linear warp and of one scaling-and-squaring step.
// Gradient semantics are those TensorFlow's autodiff gives the reference graph
// (SURVEY.md Appendix A.1): floor() has zero gradient, clip_by_value passes the gradient where
// 0 <= loc <= max (inclusive), gather back-propagates as a scatter-add.
//   d out / d loc_d   = inb_d * sum_{corners} sign_d(corner) * (prod of the other two weights) * img[corner]
//   d out / d img[k] += w_k
#include <stdlib.h>

#include "dfm_common.cuh"

namespace dfm {

struct TriG {
    uint32_t off[8];
    float w[8];
    float gx[8], gy[8], gz[8];   // d w_k / d loc_{x,y,z}
};

__device__ __forceinline__ void tri_setup_grad(float lx, float ly, float lz, int X, int Y, int Z, TriG &t) {
    const Axis ax = axis_linear(lx, (float)(X - 1));
    const Axis ay = axis_linear(ly, (float)(Y - 1));
    const Axis az = axis_linear(lz, (float)(Z - 1));
    const uint32_t YZ = (uint32_t)Y * (uint32_t)Z;
    const int ix[2] = {ax.i0, ax.i1}, iy[2] = {ay.i0, ay.i1}, iz[2] = {az.i0, az.i1};
    const float wx[2] = {ax.w0, ax.w1}, wy[2] = {ay.w0, ay.w1}, wz[2] = {az.w0, az.w1};
    const float sx[2] = {-ax.inb, ax.inb}, sy[2] = {-ay.inb, ay.inb}, sz[2] = {-az.inb, az.inb};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int a = k >> 2, b = (k >> 1) & 1, c = k & 1;
        t.off[k] = (uint32_t)ix[a] * YZ + (uint32_t)iy[b] * (uint32_t)Z + (uint32_t)iz[c];
        t.w[k] = wx[a] * wy[b] * wz[c];
        t.gx[k] = sx[a] * wy[b] * wz[c];
        t.gy[k] = wx[a] * sy[b] * wz[c];
        t.gz[k] = wx[a] * wy[b] * sz[c];
    }
}

// ---------------------------------------------------------------------------------------
// SpatialTransformer backward.  One thread per output voxel, all channels.
// ---------------------------------------------------------------------------------------
template <bool FIELD_CL, bool GFIELD_CL, bool IMG_CL>
__global__ void __launch_bounds__(256)
k_warp_bwd(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ field,
           float *__restrict__ gimg, float *__restrict__ gfield, int C, int Xi, int Yi, int Zi, int X, int Y,
           int Z, int has_fill, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zdiv);
    const uint32_t z = p - y * zdiv.d;
    const uint32_t x = blockIdx.y;
    const size_t N = (size_t)X * Y * Z, Ni = (size_t)Xi * Yi * Zi;
    const size_t vox = ((size_t)x * Y + y) * Z + z;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const float *ib = img + (size_t)blockIdx.z * C * Ni;
    const float *gb = gout + (size_t)blockIdx.z * C * N;
    float u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) u[c] = FIELD_CL ? __ldg(fb + vox * 3 + c) : __ldg(fb + c * N + vox);
    const float lx = __fadd_rn((float)x, u[0]), ly = __fadd_rn((float)y, u[1]), lz = __fadd_rn((float)z, u[2]);
    TriG t;
    tri_setup_grad(lx, ly, lz, Xi, Yi, Zi, t);
    const bool dead = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (!dead) {
        for (int c = 0; c < C; ++c) {
            const float g = IMG_CL ? __ldg(gb + vox * C + c) : __ldg(gb + (size_t)c * N + vox);
            if (gfield) {
                float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float v = IMG_CL ? __ldg(ib + (size_t)t.off[k] * C + c) : __ldg(ib + (size_t)c * Ni + t.off[k]);
                    ax = fmaf(t.gx[k], v, ax);
                    ay = fmaf(t.gy[k], v, ay);
                    az = fmaf(t.gz[k], v, az);
                }
                gx = fmaf(g, ax, gx); gy = fmaf(g, ay, gy); gz = fmaf(g, az, gz);
            }
            if (gimg) {
                float *go = gimg + (size_t)blockIdx.z * C * Ni;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float a = t.w[k] * g;
                    if (a != 0.f) atomicAdd(IMG_CL ? go + (size_t)t.off[k] * C + c : go + (size_t)c * Ni + t.off[k], a);
                }
            }
        }
    }
    if (gfield) {
        float *gf = gfield + (size_t)blockIdx.z * 3 * N;
        if (GFIELD_CL) { gf[vox * 3] = gx; gf[vox * 3 + 1] = gy; gf[vox * 3 + 2] = gz; }
        else { gf[vox] = gx; gf[N + vox] = gy; gf[2 * N + vox] = gz; }
    }
}

// ---------------------------------------------------------------------------------------
// One scaling-and-squaring step backward (planar).  v' = v + interp(v, p + v):
//   gv[p]      += s * (g[p] + sum_c g_c[p] * d interp_c / d loc)       (own + location path)
//   gv_c[k]    += s * w_k * g_c[p]                                      (volume path: adjoint of the gather)
// Two kernels share the work PER BATCH ITEM, decided on the device from a bound of the item's displacements
// (max|v_0| measured by the forward pass, doubled per step):
//   * k_ss_step_bwd_gather (|v| < 1 voxel: the early steps): NO atomics.  The volume path is evaluated as a gather
//     over the inverse neighbourhood: a source p can only touch target t if |t - p| <= 1 on every axis, and its
//     weight on t is the product of three tents max(0, 1 - |t_d - clip(p_d + v_d)|) -- exactly the corner weight
//     of the forward pass.  A thread owns a (y, z) column and marches x: the 9 sources of a plane feed three
//     running accumulators (targets x-1, x, x+1), so a voxel costs 9 x 6 loads (all L1 hits: neighbouring
//     threads read the same lines) instead of 24 float atomics, and the result is deterministic.
//   * k_ss_step_bwd (any displacement: the last steps): scatter with red.global.add, items of the other kind skipped.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool item_is_small(const float *bound, float bscale, int b) {
    return bound != nullptr && __ldg(bound + b) * bscale < 1.f;          // false for NaN
}

__global__ void __launch_bounds__(256)
k_ss_step_bwd(const float *__restrict__ g, const float *__restrict__ v, float *__restrict__ gv, int X, int Y,
              int Z, float s, FastDiv zdiv, uint32_t plane_items, const float *__restrict__ bound, float bscale) {
    if (item_is_small(bound, bscale, blockIdx.z)) return;                // served by the gather kernel
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zdiv);
    const uint32_t z = p - y * zdiv.d;
    const uint32_t x = blockIdx.y;
    const size_t N = (size_t)X * Y * Z;
    const size_t vox = ((size_t)x * Y + y) * Z + z;
    const float *vb = v + (size_t)blockIdx.z * 3 * N;
    const float *gb = g + (size_t)blockIdx.z * 3 * N;
    float *ob = gv + (size_t)blockIdx.z * 3 * N;
    const float v0 = __ldg(vb + vox), v1 = __ldg(vb + N + vox), v2 = __ldg(vb + 2 * N + vox);
    const float gg[3] = {__ldg(gb + vox), __ldg(gb + N + vox), __ldg(gb + 2 * N + vox)};
    TriG t;
    tri_setup_grad(__fadd_rn((float)x, v0), __fadd_rn((float)y, v1), __fadd_rn((float)z, v2), X, Y, Z, t);
    float gx = gg[0], gy = gg[1], gz = gg[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float val = __ldg(vb + c * N + t.off[k]);
            ax = fmaf(t.gx[k], val, ax);
            ay = fmaf(t.gy[k], val, ay);
            az = fmaf(t.gz[k], val, az);
            const float a = s * t.w[k] * gg[c];
            if (a != 0.f) atomicAdd(ob + c * N + t.off[k], a);
        }
        gx = fmaf(gg[c], ax, gx); gy = fmaf(gg[c], ay, gy); gz = fmaf(gg[c], az, gz);
    }
    atomicAdd(ob + vox, s * gx);
    atomicAdd(ob + N + vox, s * gy);
    atomicAdd(ob + 2 * N + vox, s * gz);
}

// zero the items the scatter kernel will accumulate into (the gather kernel overwrites its own)
__global__ void __launch_bounds__(256)
k_zero_large_items(float4 *__restrict__ gv, size_t n4_per_item, const float *__restrict__ bound, float bscale) {
    if (item_is_small(bound, bscale, blockIdx.y)) return;
    float4 *p = gv + (size_t)blockIdx.y * n4_per_item;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n4_per_item; i += (size_t)gridDim.x * 256ull) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

constexpr int GTY = 2;            // rows per CTA of the gather kernel
template <int NZW>
__global__ void __launch_bounds__(32 * NZW * GTY, 768 / (32 * NZW * GTY))
k_ss_step_bwd_gather(const float *__restrict__ g, const float *__restrict__ v, float *__restrict__ gv, int X, int Y, int Z,
                     float s, int seglen, const float *__restrict__ bound, float bscale) {
    const int b = blockIdx.z;
    if (!item_is_small(bound, bscale, b)) return;                        // served by the scatter kernel
    const int z = threadIdx.x, y = blockIdx.y * GTY + threadIdx.y;
    if (z >= Z || y >= Y) return;
    const int xs = blockIdx.x * seglen, xe = min(xs + seglen, X);
    const uint32_t N = (uint32_t)X * Y * Z, XS = (uint32_t)Y * Z;
    const float *vb = v + (size_t)b * 3 * N, *gb = g + (size_t)b * 3 * N;
    float *ob = gv + (size_t)b * 3 * N;
    const float mxf = (float)(X - 1), myf = (float)(Y - 1), mzf = (float)(Z - 1);
    const float fy = (float)y, fz = (float)z;
    // neighbours past the volume edge are read at the clamped position with weight 0 (branch-free: every load of a
    // plane can be in flight together)
    int yy[3], zz[3];
    float my[3], mz[3], fyy[3], fzz[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        yy[d] = min(max(y + d - 1, 0), Y - 1); my[d] = (y + d - 1 == yy[d]) ? 1.f : 0.f; fyy[d] = (float)yy[d];
        zz[d] = min(max(z + d - 1, 0), Z - 1); mz[d] = (z + d - 1 == zz[d]) ? 1.f : 0.f; fzz[d] = (float)zz[d];
    }
    float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f}, a2[3] = {0.f, 0.f, 0.f};     // targets sp - 1, sp, sp + 1
    for (int sp = xs - 1; sp <= xe; ++sp) {
        if (sp >= 0 && sp < X) {
            const float fsp = (float)sp;
            const uint32_t row0 = (uint32_t)sp * XS;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    const uint32_t idx = row0 + (uint32_t)yy[dy] * Z + (uint32_t)zz[dz];
                    // e_d = clip(p_d + v_d) - p_d lies in (-1, 1); a target one voxel below / at / above the source
                    // along an axis has the tent weight max(0, -e) / 1 - |e| / max(0, e)
                    const float ex = fminf(fmaxf(__fadd_rn(fsp, __ldg(vb + idx)), 0.f), mxf) - fsp;
                    const float ey = fminf(fmaxf(__fadd_rn(fyy[dy], __ldg(vb + N + idx)), 0.f), myf) - fyy[dy];
                    const float ez = fminf(fmaxf(__fadd_rn(fzz[dz], __ldg(vb + 2 * (size_t)N + idx)), 0.f), mzf) - fzz[dz];
                    // this thread's column sits at offset (1 - dy, 1 - dz) from the source
                    const float wy = dy == 0 ? fmaxf(0.f, ey) : dy == 1 ? 1.f - fabsf(ey) : fmaxf(0.f, -ey);
                    const float wz = dz == 0 ? fmaxf(0.f, ez) : dz == 1 ? 1.f - fabsf(ez) : fmaxf(0.f, -ez);
                    const float wyz = (my[dy] * mz[dz]) * (wy * wz);
                    // on a smooth field the sign pattern of (ey, ez) is the same across a warp, so about half of the
                    // eight neighbours have zero weight for every lane: skip their gradient loads and the 9 FMAs
                    if (__any_sync(__activemask(), wyz != 0.f)) {
                        const float t0 = wyz * fmaxf(0.f, -ex), t1 = wyz * (1.f - fabsf(ex)), t2 = wyz * fmaxf(0.f, ex);
                        const float g0 = __ldg(gb + idx), g1 = __ldg(gb + N + idx), g2 = __ldg(gb + 2 * (size_t)N + idx);
                        a0[0] = fmaf(g0, t0, a0[0]); a0[1] = fmaf(g1, t0, a0[1]); a0[2] = fmaf(g2, t0, a0[2]);
                        a1[0] = fmaf(g0, t1, a1[0]); a1[1] = fmaf(g1, t1, a1[1]); a1[2] = fmaf(g2, t1, a1[2]);
                        a2[0] = fmaf(g0, t2, a2[0]); a2[1] = fmaf(g1, t2, a2[1]); a2[2] = fmaf(g2, t2, a2[2]);
                    }
                }
            }
            if (sp >= xs && sp < xe) {                                  // own + location path of voxel (sp, y, z)
                const uint32_t vox = row0 + (uint32_t)y * Z + (uint32_t)z;
                const float gg[3] = {__ldg(gb + vox), __ldg(gb + N + vox), __ldg(gb + 2 * (size_t)N + vox)};
                TriG t;
                tri_setup_grad(__fadd_rn(fsp, __ldg(vb + vox)), __fadd_rn(fy, __ldg(vb + N + vox)),
                               __fadd_rn(fz, __ldg(vb + 2 * (size_t)N + vox)), X, Y, Z, t);
                float gx = gg[0], gy = gg[1], gz = gg[2];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float val = __ldg(vb + (size_t)c * N + t.off[k]);
                        ax = fmaf(t.gx[k], val, ax); ay = fmaf(t.gy[k], val, ay); az = fmaf(t.gz[k], val, az);
                    }
                    gx = fmaf(gg[c], ax, gx); gy = fmaf(gg[c], ay, gy); gz = fmaf(gg[c], az, gz);
                }
                a1[0] += gx; a1[1] += gy; a1[2] += gz;
            }
        }
        const int t = sp - 1;                                           // every source plane of target t has been seen
        if (t >= xs && t < xe) {
            const uint32_t o = (uint32_t)t * XS + (uint32_t)y * Z + (uint32_t)z;
            ob[o] = s * a0[0]; ob[N + o] = s * a0[1]; ob[2 * (size_t)N + o] = s * a0[2];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { a0[c] = a1[c]; a1[c] = a2[c]; a2[c] = 0.f; }
    }
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_warp_bwd(const float *gout, const float *img, const float *field, float *gimg, float *gfield,
                            int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill,
                            unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL,
                "dfm_warp_bwd: bad shape");
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_warp_bwd: B and X must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31) && (uint64_t)Xi * Yi * Zi < (1ull << 31), DFM_EINVAL,
                "dfm_warp_bwd: volume too large");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_warp_bwd: Y*Z*Z must be < 2^32");
    if (B == 0 || (!gimg && !gfield)) return DFM_OK;
    DFM_REQUIRE(gout && img && field, DFM_EINVAL, "dfm_warp_bwd: null pointer");
    if (C == 1) flags &= ~DFM_IMG_CL;
    cudaStream_t st = (cudaStream_t)stream;
    if (C > 1 && (flags & DFM_IMG_CL)) {      // channels-last multi-channel: lanes over channels
        int rc = launch_warp_cl_bwd(gout, img, field, gimg, gfield, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, flags, st);
        if (rc != DFM_EUNSUPPORTED) return rc;
    }
    if (!gimg && gfield && flags == 0u) {     // d/dfield only, all planar: TMA channel ring, no atomics
        int rc = launch_warp_mc_bwd_field(gout, img, field, gfield, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, st);
        if (rc != DFM_EUNSUPPORTED) return rc;
    }
    const uint32_t plane = (uint32_t)Y * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(Z);
    const int key = ((flags & DFM_FIELD_IN_CL) ? 4 : 0) | ((flags & DFM_FIELD_OUT_CL) ? 2 : 0) | ((flags & DFM_IMG_CL) ? 1 : 0);
#define DFM_GO(F, G, I) k_warp_bwd<F, G, I><<<grid, block, 0, st>>>(gout, img, field, gimg, gfield, C, Xi, Yi, Zi, X, Y, Z, has_fill, fd, plane)
    switch (key) {
        case 0: DFM_GO(false, false, false); break;
        case 1: DFM_GO(false, false, true); break;
        case 2: DFM_GO(false, true, false); break;
        case 3: DFM_GO(false, true, true); break;
        case 4: DFM_GO(true, false, false); break;
        case 5: DFM_GO(true, false, true); break;
        case 6: DFM_GO(true, true, false); break;
        default: DFM_GO(true, true, true); break;
    }
#undef DFM_GO
    return check_launch("dfm_warp_bwd");
}

// bound (nullable, device, one float per item) * bscale bounds |v| of the item
static int ss_step_bwd(const float *g, const float *v, float *gv, int B, int X, int Y, int Z, float scale, const float *bound,
                       float bscale, cudaStream_t st) {
    const size_t per_item = (size_t)3 * X * Y * Z;
    static const bool no_gather = getenv("DFM_NO_BWD_GATHER") != nullptr;          // tuning / testing aid
    const bool gather = bound && !no_gather && Z <= 128 && per_item % 4 == 0 && aligned16(gv);
    const uint32_t plane = (uint32_t)Y * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    if (!gather) {
        cudaError_t e = cudaMemsetAsync(gv, 0, (size_t)B * per_item * sizeof(float), st);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_ss_step_bwd: %s", cudaGetErrorString(e));
        k_ss_step_bwd<<<grid, block, 0, st>>>(g, v, gv, X, Y, Z, scale, make_fastdiv(Z), plane, nullptr, 0.f);
        return check_launch("dfm_ss_step_bwd");
    }
    k_zero_large_items<<<dim3(64, B), 256, 0, st>>>(reinterpret_cast<float4 *>(gv), per_item / 4, bound, bscale);
    // x segment per thread: long enough to amortise the two halo planes, short enough to fill the machine
    // (a thread is one (y, z) column of one segment; aim at >= 32 warps per SM)
    const int nzw = (Z + 31) / 32;
    const long long cols = (long long)B * ((Y + GTY - 1) / GTY) * GTY * nzw;      // warps per segment layer
    int seglen = 16;
    while (seglen > 4 && cols * ((X + seglen - 1) / seglen) < 148ll * 32) seglen /= 2;
    dim3 ggrid((X + seglen - 1) / seglen, (Y + GTY - 1) / GTY, B), gblock(32 * nzw, GTY);
    switch (nzw) {
        case 1: k_ss_step_bwd_gather<1><<<ggrid, gblock, 0, st>>>(g, v, gv, X, Y, Z, scale, seglen, bound, bscale); break;
        case 2: k_ss_step_bwd_gather<2><<<ggrid, gblock, 0, st>>>(g, v, gv, X, Y, Z, scale, seglen, bound, bscale); break;
        case 3: k_ss_step_bwd_gather<3><<<ggrid, gblock, 0, st>>>(g, v, gv, X, Y, Z, scale, seglen, bound, bscale); break;
        default: k_ss_step_bwd_gather<4><<<ggrid, gblock, 0, st>>>(g, v, gv, X, Y, Z, scale, seglen, bound, bscale); break;
    }
    k_ss_step_bwd<<<grid, block, 0, st>>>(g, v, gv, X, Y, Z, scale, make_fastdiv(Z), plane, bound, bscale);
    return check_launch("dfm_ss_step_bwd");
}

static int ss_step_bwd_checked(const char *who, const float *g, const float *v, float *gv, int B, int X, int Y, int Z, float scale,
                               const float *bound, float bscale, void *stream) {
    DFM_REQUIRE(B >= 0 && X >= 1 && Y >= 1 && Z >= 1 && B <= 65535 && X <= 65535, DFM_EINVAL, "%s: bad shape", who);
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31), DFM_EINVAL, "%s: volume too large", who);
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "%s: Y*Z*Z must be < 2^32", who);
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(g && v && gv && gv != g && gv != v, DFM_EINVAL, "%s: null or aliased pointer", who);
    return ss_step_bwd(g, v, gv, B, X, Y, Z, scale, bound, bscale, (cudaStream_t)stream);
}

extern "C" int dfm_ss_step_bwd(const float *g, const float *v, float *gv, int B, int X, int Y, int Z, float scale,
                               void *stream) {
    return ss_step_bwd_checked("dfm_ss_step_bwd", g, v, gv, B, X, Y, Z, scale, nullptr, 0.f, stream);
}

extern "C" int dfm_ss_step_bwd_bounded(const float *g, const float *v, float *gv, const float *bound, float bscale, int B, int X,
                                       int Y, int Z, float scale, void *stream) {
    return ss_step_bwd_checked("dfm_ss_step_bwd_bounded", g, v, gv, B, X, Y, Z, scale, bound, bscale, stream);
}

extern "C" int dfm_vecint_bwd(const float *gout, const float *saved, float *gsvf, float *scratch, int B, int X,
                              int Y, int Z, int nsteps, void *stream) {
    DFM_REQUIRE(nsteps >= 0 && nsteps <= 30, DFM_EINVAL, "dfm_vecint_bwd: nsteps %d", nsteps);
    DFM_REQUIRE(gout && gsvf, DFM_EINVAL, "dfm_vecint_bwd: null pointer");
    if (B == 0) return DFM_OK;
    const size_t n = (size_t)B * 3 * X * Y * Z;
    if (nsteps == 0) {
        cudaError_t e = cudaMemcpyAsync(gsvf, gout, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_vecint_bwd: %s", cudaGetErrorString(e));
        return DFM_OK;
    }
    DFM_REQUIRE(saved && (nsteps == 1 || scratch), DFM_EINVAL, "dfm_vecint_bwd: saved steps / scratch missing");
    const float *g = gout;
    for (int k = nsteps - 1; k >= 0; --k) {
        float *dst = (k == 0) ? gsvf : scratch + (size_t)(k & 1) * n;
        const float s = (k == 0) ? ldexpf(1.f, -nsteps) : 1.f;
        // saved ends with the forward pass's measured max|v_0| per item; |v_k| <= 2^k max|v_0|
        int rc = ss_step_bwd_checked("dfm_vecint_bwd", g, saved + (size_t)k * n, dst, B, X, Y, Z, s, saved + (size_t)nsteps * n,
                                     ldexpf(1.f, k), stream);
        if (rc) return rc;
        g = dst;
    }
    return DFM_OK;
}
