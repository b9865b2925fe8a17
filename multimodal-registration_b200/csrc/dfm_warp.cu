// SpatialTransformer forward: out[b,c,p] = interp(img[b,c], p + field[b,:,p]).
// One thread owns VEC consecutive z voxels: the field is read once (128-bit loads when planar),
// corner offsets/weights are built once per voxel and reused by every channel.
#include "dfm_common.cuh"

namespace dfm {

template <typename T, int VEC>
struct VecIO;
template <>
struct VecIO<float, 4> {
    static __device__ __forceinline__ void store(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct VecIO<uint32_t, 4> {
    static __device__ __forceinline__ void store(uint32_t *p, const uint32_t (&v)[4]) {
        *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]);
    }
};
template <typename T>
struct VecIO<T, 1> {
    static __device__ __forceinline__ void store(T *p, const T (&v)[1]) { *p = v[0]; }
};

// load the VEC field vectors owned by this thread
template <int VEC, bool FIELD_CL>
__device__ __forceinline__ void load_field(const float *fb, size_t N, size_t vox, float (&u)[3][VEC]) {
    if (FIELD_CL) {
        float a[3 * VEC];
        if (VEC == 4) {
            const float4 *q = reinterpret_cast<const float4 *>(fb + vox * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 t = __ldg(q + k);
                a[4 * k] = t.x; a[4 * k + 1] = t.y; a[4 * k + 2] = t.z; a[4 * k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) a[k] = __ldg(fb + vox * 3 + k);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) u[c][i] = a[i * 3 + c];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (VEC == 4) {
                float4 t = __ldg(reinterpret_cast<const float4 *>(fb + c * N + vox));
                u[c][0] = t.x; u[c][1] = t.y; u[c][2] = t.z; u[c][3] = t.w;
            } else {
                u[c][0] = __ldg(fb + c * N + vox);
            }
        }
    }
}

// ------------------------------- linear ------------------------------------------------
template <int VEC, bool FIELD_CL, bool IMG_CL>
__global__ void __launch_bounds__(256)
k_warp_linear(const float *__restrict__ img, const float *__restrict__ field, float *__restrict__ out,
              int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, float fill,
              int abs_loc, FastDiv zvdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zvdiv);
    const uint32_t z = (p - y * zvdiv.d) * VEC;
    const uint32_t x = blockIdx.y;
    const size_t N = (size_t)X * Y * Z, Ni = (size_t)Xi * Yi * Zi;
    const size_t vox = ((size_t)x * Y + y) * Z + z;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const float *ib = img + (size_t)blockIdx.z * C * Ni;
    float *ob = out + (size_t)blockIdx.z * C * N;

    float u[3][VEC];
    load_field<VEC, FIELD_CL>(fb, N, vox, u);

    uint32_t off[VEC][8];
    float w[VEC][8];
    bool oob[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float lx = abs_loc ? u[0][i] : __fadd_rn((float)x, u[0][i]);
        const float ly = abs_loc ? u[1][i] : __fadd_rn((float)y, u[1][i]);
        const float lz = abs_loc ? u[2][i] : __fadd_rn((float)(z + i), u[2][i]);
        tri_setup(lx, ly, lz, Xi, Yi, Zi, off[i], w[i]);
        oob[i] = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
    }

    for (int c = 0; c < C; ++c) {
        float r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float val[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                val[k] = IMG_CL ? __ldg(ib + (size_t)off[i][k] * C + c) : __ldg(ib + (size_t)c * Ni + off[i][k]);
            float acc = tri_accumulate(w[i], val);
            r[i] = oob[i] ? fill : acc;
        }
        if (IMG_CL) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) ob[(vox + i) * C + c] = r[i];
        } else {
            VecIO<float, VEC>::store(ob + (size_t)c * N + vox, r);
        }
    }
}

// ------------------------------- nearest -----------------------------------------------
template <typename T, int VEC, bool FIELD_CL, bool IMG_CL>
__global__ void __launch_bounds__(256)
k_warp_nearest(const T *__restrict__ img, const float *__restrict__ field, T *__restrict__ out,
               int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, T fill,
               int abs_loc, FastDiv zvdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t y = fast_div(p, zvdiv);
    const uint32_t z = (p - y * zvdiv.d) * VEC;
    const uint32_t x = blockIdx.y;
    const size_t N = (size_t)X * Y * Z, Ni = (size_t)Xi * Yi * Zi;
    const size_t vox = ((size_t)x * Y + y) * Z + z;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const T *ib = img + (size_t)blockIdx.z * C * Ni;
    T *ob = out + (size_t)blockIdx.z * C * N;

    float u[3][VEC];
    load_field<VEC, FIELD_CL>(fb, N, vox, u);

    uint32_t off[VEC];
    bool oob[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float lx = abs_loc ? u[0][i] : __fadd_rn((float)x, u[0][i]);
        const float ly = abs_loc ? u[1][i] : __fadd_rn((float)y, u[1][i]);
        const float lz = abs_loc ? u[2][i] : __fadd_rn((float)(z + i), u[2][i]);
        off[i] = ((uint32_t)axis_nearest(lx, Xi - 1) * Yi + axis_nearest(ly, Yi - 1)) * Zi +
                 axis_nearest(lz, Zi - 1);
        oob[i] = has_fill && oob3(lx, ly, lz, Xi, Yi, Zi);
    }
    for (int c = 0; c < C; ++c) {
        T r[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            T val = IMG_CL ? ib[(size_t)off[i] * C + c] : ib[(size_t)c * Ni + off[i]];
            r[i] = oob[i] ? fill : val;
        }
        if (IMG_CL) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) ob[(vox + i) * C + c] = r[i];
        } else {
            VecIO<T, VEC>::store(ob + (size_t)c * N + vox, r);
        }
    }
}

template <int VEC>
static int launch_linear(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi,
                         int Zi, int X, int Y, int Z, int has_fill, float fill, unsigned flags,
                         cudaStream_t st) {
    const uint32_t zv = Z / VEC, plane = (uint32_t)Y * zv;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(zv);
    const bool fcl = flags & DFM_FIELD_IN_CL, icl = flags & DFM_IMG_CL;
    const int abs_loc = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
#define DFM_GO(F, I) k_warp_linear<VEC, F, I><<<grid, block, 0, st>>>( \
        img, field, out, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, abs_loc, fd, plane)
    if (fcl) { if (icl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (icl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_warp_fwd(linear)");
}

template <typename T, int VEC>
static int launch_nearest(const void *img, const float *field, void *out, int B, int C, int Xi, int Yi,
                          int Zi, int X, int Y, int Z, int has_fill, uint64_t fill_bits, unsigned flags,
                          cudaStream_t st) {
    const uint32_t zv = Z / VEC, plane = (uint32_t)Y * zv;
    dim3 grid((plane + 255) / 256, X, B), block(256);
    FastDiv fd = make_fastdiv(zv);
    const bool fcl = flags & DFM_FIELD_IN_CL, icl = flags & DFM_IMG_CL;
    const int abs_loc = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
    const T fill = (T)fill_bits;
#define DFM_GO(F, I) k_warp_nearest<T, VEC, F, I><<<grid, block, 0, st>>>( \
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
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31) && (uint64_t)Xi * Yi * Zi < (1ull << 31), DFM_EINVAL,
                "dfm_warp_fwd: volume too large (>= 2^31 voxels)");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_warp_fwd: Y*Z*Z must be < 2^32");
    DFM_REQUIRE(img && field && out, DFM_EINVAL, "dfm_warp_fwd: null pointer");
    DFM_REQUIRE(img != out, DFM_EINVAL, "dfm_warp_fwd: out must not alias img");
    if (B == 0) return DFM_OK;
    if (C == 1) flags &= ~DFM_IMG_CL;   // one channel: both layouts coincide
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec4 = (Z % 4 == 0) && aligned16(field) && aligned16(out);
    if (interp == DFM_LINEAR) {
        DFM_REQUIRE(elem_size == 4, DFM_EINVAL, "dfm_warp_fwd: linear interpolation needs fp32 (elem_size 4), got %d", elem_size);
        return vec4 ? launch_linear<4>((const float *)img, field, (float *)out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st)
                    : launch_linear<1>((const float *)img, field, (float *)out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
    }
    DFM_REQUIRE(interp == DFM_NEAREST, DFM_EINVAL, "dfm_warp_fwd: interp %d", interp);
    switch (elem_size) {
        case 4:
            return vec4 ? launch_nearest<uint32_t, 4>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st)
                        : launch_nearest<uint32_t, 1>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 1: return launch_nearest<uint8_t, 1>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 2: return launch_nearest<uint16_t, 1>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        case 8: return launch_nearest<uint64_t, 1>(img, field, out, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill_bits, flags, st);
        default: return fail(DFM_EINVAL, "dfm_warp_fwd: elem_size %d not in {1,2,4,8}", elem_size);
    }
}
