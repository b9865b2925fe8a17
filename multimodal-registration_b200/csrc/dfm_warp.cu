// SpatialTransformer forward: out[b,c,p] = interp(img[b,c], p + field[b,:,p]).
// Corner offsets/weights are built once per voxel and reused by every channel.
#include <stdlib.h>

#include "dfm_common.cuh"

namespace dfm {

// one field vector (planar or channels-last)
template <bool FIELD_CL>
__device__ __forceinline__ void load_field1(const float *fb, uint32_t N, uint32_t vox, float &u0, float &u1, float &u2) {
    if (FIELD_CL) {
        u0 = __ldg(fb + (size_t)vox * 3); u1 = __ldg(fb + (size_t)vox * 3 + 1); u2 = __ldg(fb + (size_t)vox * 3 + 2);
    } else {
        u0 = __ldg(fb + vox); u1 = __ldg(fb + N + vox); u2 = __ldg(fb + 2 * (size_t)N + vox);
    }
}

// ------------------------------- linear ------------------------------------------------
// Lanes own 32 consecutive z of one row (every gather touches 1-3 lines), a thread owns ROWS
// consecutive rows (independent chains -> ILP).  ONE_CH specialises the C = 1 image warp.
template <int ROWS, bool ONE_CH, bool FIELD_CL, bool IMG_CL>
__global__ void __launch_bounds__(256, 4)
k_warp_linear(const float *__restrict__ img, const float *__restrict__ field, float *__restrict__ out,
              int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, float fill,
              int abs_loc, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t yy = fast_div(p, zdiv);
    const uint32_t z = p - yy * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const float *ib = img + (size_t)blockIdx.z * C * Ni;
    float *ob = out + (size_t)blockIdx.z * C * N;
    const float fx = (float)x, fz = (float)z;
    const bool fast = Xi >= 2 && Yi >= 2 && Zi >= 2;                              // uniform
    const uint32_t es = (!ONE_CH && IMG_CL) ? (uint32_t)C : 1u;                  // element stride
    const uint32_t gy = es * Zi, gx = gy * Yi;

#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t y = yy * ROWS + r;
        if (y >= (uint32_t)Y) break;
        const uint32_t vox = (x * Y + y) * Z + z;
        float u0, u1, u2;
        load_field1<FIELD_CL>(fb, N, vox, u0, u1, u2);
        const float lx = abs_loc ? u0 : __fadd_rn(fx, u0);
        const float ly = abs_loc ? u1 : __fadd_rn((float)y, u1);
        const float lz = abs_loc ? u2 : __fadd_rn(fz, u2);
        float w[8];
        const bool oob = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
        if (fast) {
            const uint32_t base = tri_setup_fast(lx, ly, lz, Xi, Yi, Zi, w);
            if (ONE_CH) {
                float val[8];
                gather8(ib + base, gy, gx, 1u, val);
                const float acc = tri_accumulate(w, val);
                ob[vox] = oob ? fill : acc;
            } else {
                const float *ic = IMG_CL ? ib + (size_t)base * C : ib + base;
                float *oc = IMG_CL ? ob + (size_t)vox * C : ob + vox;
                for (int c = 0; c < C; ++c) {
                    float val[8];
                    gather8(ic, gy, gx, es, val);
                    const float acc = tri_accumulate(w, val);
                    *oc = oob ? fill : acc;
                    ic += IMG_CL ? 1 : Ni;
                    oc += IMG_CL ? 1 : N;
                }
            }
        } else {
            uint32_t off[8];
            tri_setup(lx, ly, lz, Xi, Yi, Zi, off, w);
            const float *ic = ib;
            float *oc = IMG_CL ? ob + (size_t)vox * C : ob + vox;
            for (int c = 0; c < C; ++c) {
                float val[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) val[k] = IMG_CL ? __ldg(ic + (size_t)off[k] * C) : __ldg(ic + off[k]);
                const float acc = tri_accumulate(w, val);
                *oc = oob ? fill : acc;
                ic += IMG_CL ? 1 : Ni;
                oc += IMG_CL ? 1 : N;
            }
        }
    }
}

// ------------------------------- nearest -----------------------------------------------
template <typename T, int ROWS, bool FIELD_CL, bool IMG_CL>
__global__ void __launch_bounds__(256, 4)
k_warp_nearest(const T *__restrict__ img, const float *__restrict__ field, T *__restrict__ out,
               int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, T fill,
               int abs_loc, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t yy = fast_div(p, zdiv);
    const uint32_t z = p - yy * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const T *ib = img + (size_t)blockIdx.z * C * Ni;
    T *ob = out + (size_t)blockIdx.z * C * N;
    const float fx = (float)x, fz = (float)z;

#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t y = yy * ROWS + r;
        if (y >= (uint32_t)Y) break;
        const uint32_t vox = (x * Y + y) * Z + z;
        float u0, u1, u2;
        load_field1<FIELD_CL>(fb, N, vox, u0, u1, u2);
        const float lx = abs_loc ? u0 : __fadd_rn(fx, u0);
        const float ly = abs_loc ? u1 : __fadd_rn((float)y, u1);
        const float lz = abs_loc ? u2 : __fadd_rn(fz, u2);
        const uint32_t off = ((uint32_t)axis_nearest(lx, Xi - 1) * Yi + axis_nearest(ly, Yi - 1)) * Zi +
                             axis_nearest(lz, Zi - 1);
        const bool oob = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
        const T *ic = IMG_CL ? ib + (size_t)off * C : ib + off;
        T *oc = IMG_CL ? ob + (size_t)vox * C : ob + vox;
        for (int c = 0; c < C; ++c) {
            const T val = *ic;
            *oc = oob ? fill : val;
            ic += IMG_CL ? 1 : Ni;
            oc += IMG_CL ? 1 : N;
        }
    }
}

// one-channel linear warp, direct gathers, written for memory-level parallelism: a thread first
// loads the field vectors of all its rows, then issues all 8*ROWS corner gathers, and only then
// forms the weights and accumulates -- the loads of a thread are all in flight together, and the
// small register footprint keeps many warps resident.  (Needs every image axis >= 2.)
template <int ROWS, bool FIELD_CL>
__global__ void __launch_bounds__(256, (ROWS == 1) ? 8 : 5)
k_warp_linear1(const float *__restrict__ img, const float *__restrict__ field, float *__restrict__ out,
               int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, float fill,
               int abs_loc, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t yy = fast_div(p, zdiv);
    const uint32_t z = p - yy * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const float *ib = img + (size_t)blockIdx.z * Ni;
    float *ob = out + (size_t)blockIdx.z * N;
    const float fx = (float)x, fz = (float)z;
    const uint32_t vox0 = (x * Y + yy * ROWS) * Z + z;
    const int nr = min(ROWS, Y - (int)(yy * ROWS));
    const uint32_t gy = (uint32_t)Zi, gx = (uint32_t)Yi * Zi;
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    float l[ROWS][3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        l[r][0] = l[r][1] = l[r][2] = 0.f;
        if (r < nr) load_field1<FIELD_CL>(fb, N, vox0 + r * Z, l[r][0], l[r][1], l[r][2]);
    }
    float val[ROWS][8];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (!abs_loc) {
            l[r][0] = __fadd_rn(fx, l[r][0]);
            l[r][1] = __fadd_rn((float)(yy * ROWS + r), l[r][1]);
            l[r][2] = __fadd_rn(fz, l[r][2]);
        }
        const uint32_t base = ((uint32_t)(axis_fast_i1(l[r][0], mxf, mxi) - 1) * Yi + (uint32_t)(axis_fast_i1(l[r][1], myf, myi) - 1)) * Zi +
                              (uint32_t)(axis_fast_i1(l[r][2], mzf, mzi) - 1);
        gather8(ib + base, gy, gx, 1u, val[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (r >= nr) break;
        float w[8];
        tri_weights(axis_fast(l[r][0], mxf, mxi), axis_fast(l[r][1], myf, myi), axis_fast(l[r][2], mzf, mzi), w);
        float acc = tri_accumulate(w, val[r]);
        if (has_fill && oob3(l[r][0], l[r][1], l[r][2], Xi, Yi, Zi)) acc = fill;
        ob[vox0 + r * Z] = acc;
    }
}

// one-channel nearest warp (label maps): latency-bound, so every thread first issues the field
// loads of all its rows, then all gathers, then the stores (memory-level parallelism), and the
// register budget allows 8 CTAs per SM
template <typename T, int ROWS, bool FIELD_CL>
__global__ void __launch_bounds__(256, 8)
k_warp_nearest1(const T *__restrict__ img, const float *__restrict__ field, T *__restrict__ out,
                int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, T fill,
                int abs_loc, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t yy = fast_div(p, zdiv);
    const uint32_t z = p - yy * zdiv.d;
    const uint32_t x = blockIdx.y;
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const T *ib = img + (size_t)blockIdx.z * Ni;
    T *ob = out + (size_t)blockIdx.z * N;
    const float fx = (float)x, fz = (float)z;
    const uint32_t vox0 = (x * Y + yy * ROWS) * Z + z;
    const int nr = min(ROWS, Y - (int)(yy * ROWS));
    float u[ROWS][3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        u[r][0] = u[r][1] = u[r][2] = 0.f;
        if (r < nr) load_field1<FIELD_CL>(fb, N, vox0 + r * Z, u[r][0], u[r][1], u[r][2]);
    }
    T val[ROWS];
    bool oob[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const float lx = abs_loc ? u[r][0] : __fadd_rn(fx, u[r][0]);
        const float ly = abs_loc ? u[r][1] : __fadd_rn((float)(yy * ROWS + r), u[r][1]);
        const float lz = abs_loc ? u[r][2] : __fadd_rn(fz, u[r][2]);
        const uint32_t off = ((uint32_t)axis_nearest(lx, Xi - 1) * Yi + axis_nearest(ly, Yi - 1)) * Zi + axis_nearest(lz, Zi - 1);
        oob[r] = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
        val[r] = (r < nr) ? ib[off] : fill;
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
        if (r < nr) ob[vox0 + r * Z] = oob[r] ? fill : val[r];
}

constexpr int WARP_ROWS = 4;

static int launch_linear(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi,
                         int Zi, int X, int Y, int Z, int has_fill, float fill, unsigned flags,
                         cudaStream_t st) {
    if (C == 1 && Xi >= 2 && Yi >= 2 && Zi >= 2) {
        static const int r1 = getenv("DFM_WARP_ROWS1") ? 1 : 0;      // tuning aid
        const int R = r1 ? 1 : 2;
        const uint32_t pl = (uint32_t)((Y + R - 1) / R) * Z;
        dim3 g((pl + 255) / 256, X, B), b(256);
        const FastDiv fz_ = make_fastdiv(Z);
        const int al = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
        const bool fc = flags & DFM_FIELD_IN_CL;
#define DFM_G1(RR, F) k_warp_linear1<RR, F><<<g, b, 0, st>>>(img, field, out, Xi, Yi, Zi, X, Y, Z, has_fill, fill, al, fz_, pl)
        if (R == 1) { if (fc) DFM_G1(1, true); else DFM_G1(1, false); }
        else        { if (fc) DFM_G1(2, true); else DFM_G1(2, false); }
#undef DFM_G1
        return check_launch("dfm_warp_fwd(linear, C=1 direct)");
    }
    const int rows = WARP_ROWS;
    const uint32_t plane = (uint32_t)((Y + rows - 1) / rows) * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(Z);
    const bool fcl = flags & DFM_FIELD_IN_CL, icl = flags & DFM_IMG_CL;
    const int abs_loc = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
#define DFM_GO(O, F, I) k_warp_linear<WARP_ROWS, O, F, I><<<grid, block, 0, st>>>( \
        img, field, out, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, abs_loc, fd, plane)
    if (C == 1) { if (fcl) DFM_GO(true, true, false); else DFM_GO(true, false, false); }
    else if (fcl) { if (icl) DFM_GO(false, true, true); else DFM_GO(false, true, false); }
    else          { if (icl) DFM_GO(false, false, true); else DFM_GO(false, false, false); }
#undef DFM_GO
    return check_launch("dfm_warp_fwd(linear)");
}

template <typename T>
static int launch_nearest(const void *img, const float *field, void *out, int B, int C, int Xi, int Yi,
                          int Zi, int X, int Y, int Z, int has_fill, uint64_t fill_bits, unsigned flags,
                          cudaStream_t st) {
    const uint32_t plane = (uint32_t)((Y + WARP_ROWS - 1) / WARP_ROWS) * Z;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(Z);
    const bool fcl = flags & DFM_FIELD_IN_CL, icl = flags & DFM_IMG_CL;
    const int abs_loc = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
    const T fill = (T)fill_bits;
    if (C == 1) {
        if (fcl) k_warp_nearest1<T, WARP_ROWS, true><<<grid, block, 0, st>>>((const T *)img, field, (T *)out, Xi, Yi, Zi, X, Y, Z, has_fill, fill, abs_loc, fd, plane);
        else k_warp_nearest1<T, WARP_ROWS, false><<<grid, block, 0, st>>>((const T *)img, field, (T *)out, Xi, Yi, Zi, X, Y, Z, has_fill, fill, abs_loc, fd, plane);
        return check_launch("dfm_warp_fwd(nearest, C=1)");
    }
#define DFM_GO(F, I) k_warp_nearest<T, WARP_ROWS, F, I><<<grid, block, 0, st>>>( \
        (const T *)img, field, (T *)out, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, abs_loc, fd, plane)
    if (fcl) { if (icl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (icl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_warp_fwd(nearest)");
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_warp_fwd(const void *img, const float *field, void *out, int B, int C, int Xi, int Yi,
                            int Zi, int X, int Y, int Z, int interp, int elem_size, int has_fill, float fill,
                            uint64_t fill_bits, unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL,
                "dfm_warp_fwd: bad shape B=%d C=%d img=(%d,%d,%d) grid=(%d,%d,%d)", B, C, Xi, Yi, Zi, X, Y, Z);
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_warp_fwd: B and X must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 30) && (uint64_t)Xi * Yi * Zi < (1ull << 30), DFM_EINVAL,
                "dfm_warp_fwd: volume too large (>= 2^30 voxels)");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_warp_fwd: Y*Z*Z must be < 2^32");
    DFM_REQUIRE(img && field && out, DFM_EINVAL, "dfm_warp_fwd: null pointer");
    DFM_REQUIRE(img != out, DFM_EINVAL, "dfm_warp_fwd: out must not alias img");
    if (B == 0) return DFM_OK;
    if (C == 1) flags &= ~DFM_IMG_CL;   // one channel: both layouts coincide
    cudaStream_t st = (cudaStream_t)stream;
    if (interp == DFM_LINEAR) {
        DFM_REQUIRE(elem_size == 4, DFM_EINVAL, "dfm_warp_fwd: linear interpolation needs fp32 (elem_size 4), got %d", elem_size);
        if (C == 1) {   // one channel: texture gathers where selected (dfm_warp_tex.cu), else the TMA-brick path (dfm_brick.cu)
            int rc = launch_warp_tex((const float *)img, field, (float *)out, B, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
            rc = launch_warp_brick((const float *)img, field, (float *)out, B, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
        }
        if (C > 1 && (flags & DFM_IMG_CL)) {   // channels-last multi-channel (the reference layout): lanes over channels
            int rc = launch_warp_cl_fwd((const float *)img, field, (float *)out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
        }
        if (C > 1 && !(flags & (DFM_IMG_CL | DFM_FIELD_IN_CL | DFM_LOC_ABSOLUTE))) {   // planar multi-channel: TMA channel ring
            int rc = launch_warp_mc_fwd((const float *)img, field, (float *)out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
        }
        return launch_linear((const float *)img, field, (float *)out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
    }
    DFM_REQUIRE(interp == DFM_NEAREST, DFM_EINVAL, "dfm_warp_fwd: interp %d", interp);
    switch (elem_size) {
        case 4: return launch_nearest<uint32_t>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 1: return launch_nearest<uint8_t>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 2: return launch_nearest<uint16_t>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 8: return launch_nearest<uint64_t>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        default: return fail(DFM_EINVAL, "dfm_warp_fwd: elem_size %d not in {1,2,4,8}", elem_size);
    }
}

extern "C" int dfm_rescale_warp_fwd(const float *img, const float *coarse, float *out, const float *cx, const float *cy,
                                    const float *cz, float *work, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh,
                                    int X, int Y, int Z, float factor, int has_fill, float fill, void *stream) {
    DFM_REQUIRE(B >= 0 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xh >= 1 && Yh >= 1 && Zh >= 1 && X >= 1 && Y >= 1 && Z >= 1,
                DFM_EINVAL, "dfm_rescale_warp_fwd: bad shape");
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_rescale_warp_fwd: B and X must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 30) && (uint64_t)Xi * Yi * Zi < (1ull << 30), DFM_EINVAL,
                "dfm_rescale_warp_fwd: volume too large (>= 2^30 voxels)");
    DFM_REQUIRE(factor >= 1.f, DFM_EINVAL, "dfm_rescale_warp_fwd: factor %g < 1 (scale-then-resize order only)", factor);
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(img && coarse && out && cx && cy && cz, DFM_EINVAL, "dfm_rescale_warp_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_rescale_warp_tex(img, coarse, out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, factor, has_fill, fill, st);
    if (rc != DFM_EUNSUPPORTED) return rc;
    rc = launch_rescale_warp(img, coarse, out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, factor, has_fill, fill, st);
    if (rc != DFM_EUNSUPPORTED) return rc;
    DFM_REQUIRE(work, DFM_EUNSUPPORTED, "dfm_rescale_warp_fwd: fused kernel not applicable to this shape and no work buffer given");
    rc = dfm_resize_fwd(coarse, work, cx, cy, cz, B, 3, Xh, Yh, Zh, X, Y, Z, factor, 1.f, DFM_LINEAR, 0u, stream);
    if (rc) return rc;
    return dfm_warp_fwd(img, work, out, B, 1, Xi, Yi, Zi, X, Y, Z, DFM_LINEAR, 4, has_fill, fill, 0, 0u, stream);
}

extern "C" int dfm_rescale_warp_nearest_fwd(const void *img, const float *coarse, void *out, const float *cx, const float *cy,
                                            const float *cz, float *work, int B, int Xi, int Yi, int Zi, int Xh, int Yh,
                                            int Zh, int X, int Y, int Z, float factor, int has_fill, uint32_t fill_bits,
                                            void *stream) {
    DFM_REQUIRE(B >= 0 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xh >= 1 && Yh >= 1 && Zh >= 1 && X >= 1 && Y >= 1 && Z >= 1,
                DFM_EINVAL, "dfm_rescale_warp_nearest_fwd: bad shape");
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_rescale_warp_nearest_fwd: B and X must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 30) && (uint64_t)Xi * Yi * Zi < (1ull << 30), DFM_EINVAL,
                "dfm_rescale_warp_nearest_fwd: volume too large (>= 2^30 voxels)");
    DFM_REQUIRE(factor >= 1.f, DFM_EINVAL, "dfm_rescale_warp_nearest_fwd: factor %g < 1 (scale-then-resize order only)", factor);
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(img && coarse && out && cx && cy && cz, DFM_EINVAL, "dfm_rescale_warp_nearest_fwd: null pointer");
    DFM_REQUIRE(img != out, DFM_EINVAL, "dfm_rescale_warp_nearest_fwd: out must not alias img");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_rescale_warp_nearest(img, coarse, out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, factor, has_fill, fill_bits, st);
    if (rc != DFM_EUNSUPPORTED) return rc;
    DFM_REQUIRE(work, DFM_EUNSUPPORTED, "dfm_rescale_warp_nearest_fwd: fused kernel not applicable to this shape and no work buffer given");
    rc = dfm_resize_fwd(coarse, work, cx, cy, cz, B, 3, Xh, Yh, Zh, X, Y, Z, factor, 1.f, DFM_LINEAR, 0u, stream);
    if (rc) return rc;
    return dfm_warp_fwd(img, work, out, B, 1, Xi, Yi, Zi, X, Y, Z, DFM_NEAREST, 4, has_fill, 0.f, (uint64_t)fill_bits, 0u, stream);
}
