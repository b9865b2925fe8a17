// Backward passes of the linear warp and of one scaling-and-squaring step.
// Gradient semantics are those TensorFlow's autodiff gives the reference graph
// (SURVEY.md Appendix A.1): floor() has zero gradient, clip_by_value passes the gradient where
// 0 <= loc <= max (inclusive), gather back-propagates as a scatter-add.
//   d out / d loc_d   = inb_d * sum_{corners} sign_d(corner) * (prod of the other two weights) * img[corner]
//   d out / d img[k] += w_k
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
//   gv_c[k]    += s * w_k * g_c[p]                                      (volume path, scatter)
// gv must be zero on entry (the API call clears it).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ss_step_bwd(const float *__restrict__ g, const float *__restrict__ v, float *__restrict__ gv, int X, int Y,
              int Z, float s, FastDiv zdiv, uint32_t plane_items) {
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

extern "C" int dfm_ss_step_bwd(const float *g, const float *v, float *gv, int B, int X, int Y, int Z, float scale,
                               void *stream) {
    DFM_REQUIRE(B >= 0 && X >= 1 && Y >= 1 && Z >= 1 && B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_ss_step_bwd: bad shape");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31), DFM_EINVAL, "dfm_ss_step_bwd: volume too large");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_ss_step_bwd: Y*Z*Z must be < 2^32");
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(g && v && gv && gv != g && gv != v, DFM_EINVAL, "dfm_ss_step_bwd: null or aliased pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)B * 3 * X * Y * Z * sizeof(float);
    cudaError_t e = cudaMemsetAsync(gv, 0, bytes, st);
    DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_ss_step_bwd: %s", cudaGetErrorString(e));
    const uint32_t plane = (uint32_t)Y * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    k_ss_step_bwd<<<grid, block, 0, st>>>(g, v, gv, X, Y, Z, scale, make_fastdiv(Z), plane);
    return check_launch("dfm_ss_step_bwd");
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
        int rc = dfm_ss_step_bwd(g, saved + (size_t)k * n, dst, B, X, Y, Z, s, stream);
        if (rc) return rc;
        g = dst;
    }
    return DFM_OK;
}
