// Jacobian-determinant map of a displacement field (eval_reg_with_jacobian.py:62-78):
// 4th-order central differences on the interior [2:-2]^3, det(I + J) in fp64, fold count and
// moments.  fp64 in registers is free under the memory roof; HBM sees the field once
// (12 B/voxel as fp32) and the determinant map once.
#include <cuda.h>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

namespace dfm {

template <typename T>
__device__ __forceinline__ double ldd(const T *p) { return (double)__ldg(p); }

// (u[-2] - 8 u[-1] + 8 u[+1] - u[+2]) / 12, evaluated left to right like the reference (:66-68)
__device__ __forceinline__ double d4(double m2, double m1, double p1, double p2) {
    double t = __dsub_rn(m2, __dmul_rn(8.0, m1));
    t = __dadd_rn(t, __dmul_rn(8.0, p1));
    t = __dsub_rn(t, p2);
    return t / 12.0;
}

template <typename Tin, typename Tout, bool IN_CL>
__global__ void __launch_bounds__(256)
k_jacdet(const Tin *__restrict__ field, Tout *__restrict__ det, double *__restrict__ partials, int X, int Y,
         int Z, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const int Zo = Z - 4, Yo = Y - 4, Xo = X - 4;
    const size_t N = (size_t)X * Y * Z;
    const Tin *fb = field + (size_t)blockIdx.z * 3 * N;
    double dval = 0.0;
    bool valid = p < plane_items;
    if (valid) {
        const uint32_t yo = fast_div(p, zdiv);
        const uint32_t zo = p - yo * zdiv.d;
        const uint32_t xo = blockIdx.y;
        const size_t ctr = ((size_t)(xo + 2) * Y + (yo + 2)) * Z + (zo + 2);
        const size_t sx = (size_t)Y * Z, sy = Z;
        double J[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Tin *q = IN_CL ? fb + c : fb + c * N;
            const size_t m = IN_CL ? 3 : 1;
            J[c][0] = d4(ldd(q + (ctr - 2 * sx) * m), ldd(q + (ctr - sx) * m), ldd(q + (ctr + sx) * m), ldd(q + (ctr + 2 * sx) * m));
            J[c][1] = d4(ldd(q + (ctr - 2 * sy) * m), ldd(q + (ctr - sy) * m), ldd(q + (ctr + sy) * m), ldd(q + (ctr + 2 * sy) * m));
            J[c][2] = d4(ldd(q + (ctr - 2) * m), ldd(q + (ctr - 1) * m), ldd(q + (ctr + 1) * m), ldd(q + (ctr + 2) * m));
        }
        J[0][0] += 1.0; J[1][1] += 1.0; J[2][2] += 1.0;
        dval = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
               J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        if (det) det[(size_t)blockIdx.z * Xo * Yo * Zo + ((size_t)xo * Yo + yo) * Zo + zo] = (Tout)dval;
    }
    if (!partials) return;
    // block reduction in a fixed order -> deterministic statistics
    double s = valid ? dval : 0.0, s2 = valid ? dval * dval : 0.0;
    double nn = (valid && !(dval > 0.0) && dval != 0.0) ? 1.0 : 0.0;   // np.count_nonzero(np.where(det > 0, 0, det)): negatives and NaN
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nn += __shfl_down_sync(0xffffffffu, nn, o);
    }
    __shared__ double sh[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = nn; sh[1][warp] = s; sh[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += sh[threadIdx.x][k];
        const size_t nblk = (size_t)gridDim.x * gridDim.y;
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        partials[((size_t)blockIdx.z * 3 + threadIdx.x) * nblk + blk] = a;
    }
}

__global__ void __launch_bounds__(256)
k_jacdet_finalize(const double *__restrict__ partials, double *__restrict__ stats, size_t nblk, double ntotal) {
    __shared__ double sh[3][256];
    const double *pb = partials + (size_t)blockIdx.x * 3 * nblk;
    for (int q = 0; q < 3; ++q) {
        double a = 0.0;
        for (size_t k = threadIdx.x; k < nblk; k += 256) a += pb[q * nblk + k];
        sh[q][threadIdx.x] = a;
    }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *st = stats + (size_t)blockIdx.x * 4;
        st[0] = sh[0][0]; st[1] = sh[1][0]; st[2] = sh[2][0]; st[3] = ntotal;
    }
}

// ---------------------------------------------------------------------------------------
// Fast path for planar fp32 fields: plane-marching tile kernel.
// A CTA owns 32 x 16 x 32 (x, y, z) determinants.  It marches along x; every step the input
// plane x+2 (16+4 rows x 32+4 columns x 3 components) is staged in a 4-slot shared-memory ring.
// A thread owns two (y, z) columns: the x stencil comes from a 5-deep register window of its own
// column, the y and z stencils from the shared plane.  The four-point stencil is evaluated in
// fp32 in difference form  ((u[-2] - u[+2]) + 8 (u[+1] - u[-1]))  -- differences of neighbouring
// samples, so the fp32 rounding is relative to the local variation of the field (~1e-7 for
// registration fields); the three 2x2 minors are fp32 (one fused rounding each), the expansion along
// the first row, the fold test and the moments are fp64.  (The all-fp64 kernel above needs 36 fp32->fp64 conversions
// per voxel, which run at 1/8 rate; it is kept for fp64 inputs and channels-last fields.)
// ---------------------------------------------------------------------------------------
constexpr int JT_X = 64, JT_Y = 16, JT_Z = 32, JP_Y = JT_Y + 4, JP_Z = JT_Z + 4, J_SLOTS = 5;

__device__ __forceinline__ float d4f(float m2, float m1, float p1, float p2) {
    return fmaf(8.f, p1 - m1, m2 - p2);            // 12 * derivative
}

// det(I + grad u) from J = 12 * (I + grad u) with the 12 still to be added on the diagonal:
// 2x2 minors in fp32 with one fused rounding each, combined and scaled in fp64
__device__ __forceinline__ double det12(float (&J)[3][3]) {
    J[0][0] += 12.f; J[1][1] += 12.f; J[2][2] += 12.f;
    const float m0 = fmaf(J[1][1], J[2][2], -J[1][2] * J[2][1]);
    const float m1 = fmaf(J[1][0], J[2][2], -J[1][2] * J[2][0]);
    const float m2 = fmaf(J[1][0], J[2][1], -J[1][1] * J[2][0]);
    return ((double)J[0][0] * (double)m0 - (double)J[0][1] * (double)m1 + (double)J[0][2] * (double)m2) * (1.0 / 1728.0);
}
__device__ __forceinline__ void store_pair(float *q, double a, double b) { *reinterpret_cast<float2 *>(q) = make_float2((float)a, (float)b); }
__device__ __forceinline__ void store_pair(double *q, double a, double b) { *reinterpret_cast<double2 *>(q) = make_double2(a, b); }

// TMA plane ring: one 4-D box {JP_Z, JP_Y, 1, 3} per input plane, J_SLOTS deep, one mbarrier per
// slot.  A slot is re-armed only after the __syncthreads() that follows its last reader.
template <typename Tout>
__global__ void __launch_bounds__(256)
k_jacdet_tiled(const __grid_constant__ CUtensorMap tmap, Tout *__restrict__ det, double *__restrict__ partials,
               int X, int Y, int Z, int nzt) {
    // every slot starts on a 128-byte boundary (TMA destination alignment): 2160 floats padded to 2176
    constexpr int SLOT_FLOATS = ((3 * JP_Y * JP_Z + 31) / 32) * 32;
    __shared__ __align__(128) float plane_raw[J_SLOTS][SLOT_FLOATS];
#define PL(slot, c, y, z) plane_raw[slot][((c) * JP_Y + (y)) * JP_Z + (z)]
    __shared__ __align__(8) uint64_t bar[J_SLOTS];
    constexpr uint32_t PLANE_BYTES = 3 * JP_Y * JP_Z * sizeof(float);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int zo0 = zt * JT_Z, yo0 = yt * JT_Y, xo0 = blockIdx.y * JT_X;
    const int Xo = X - 4, Yo = Y - 4, Zo = Z - 4;
    const int nxo = min(JT_X, Xo - xo0);                     // output planes of this CTA
    const int np = nxo + 4;                                  // input planes
    // a thread owns TWO z-adjacent outputs (z even) of one row: every stencil read is an aligned
    // 8-byte shared load (LDS.64) and the z stencil shares its taps between the two outputs
    const int row = warp * 2 + (lane >> 4), zp = (lane & 15) * 2;
    const int zo = zo0 + zp, yo = yo0 + row;
    const bool ok0 = yo < Yo && zo < Zo, ok1 = yo < Yo && (zo + 1) < Zo;
    const bool pair_store = ok1 && !(Zo & 1);                // 8/16-byte aligned pair store

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < J_SLOTS; ++k) mbar_init(&bar[k], 1);
    }
    __syncthreads();
    auto issue = [&](int p) {                                // thread 0 only
        mbar_expect_tx(&bar[p % J_SLOTS], PLANE_BYTES);
        tma_load_4d(&plane_raw[p % J_SLOTS][0], &tmap, &bar[p % J_SLOTS], zo0, yo0, xo0 + p, (int)blockIdx.z * 3);
    };
    if (threadIdx.x == 0)
        for (int p = 0; p < min(np, J_SLOTS - 2); ++p) issue(p);

    float win[3][2][5];                                      // [component][z / z+1][plane ring]
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 5; ++k) win[c][r][k] = 0.f;
    double s = 0.0, s2 = 0.0, nn = 0.0;
    Tout *dp = det ? det + (size_t)blockIdx.z * Xo * Yo * Zo + ((size_t)xo0 * Yo + yo) * Zo + zo : nullptr;
    const size_t dplane = (size_t)Yo * Zo;
#define PL2(slot, c, y, z) (*reinterpret_cast<const float2 *>(&PL(slot, c, y, z)))

    // one step: plane p arrives, its centres enter the register window at position (p % 5), and if
    // p >= 4 the determinants of output plane p - 4 (centre plane p - 2) are produced.  The window is
    // indexed with compile-time rotations K = p % 5, so it never moves between registers.
#define JAC_STEP(K)                                                                                                   \
    if (p < np) {                                                                                                     \
        const int slot = p % J_SLOTS;                                                                                 \
        mbar_wait(&bar[slot], (uint32_t)((p / J_SLOTS) & 1));                                                         \
        _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                                               \
            const float2 t = PL2(slot, c, row + 2, zp + 2);                                                           \
            win[c][0][K] = t.x; win[c][1][K] = t.y;                                                                   \
        }                                                                                                             \
        if (p >= 4) {                                                                                                 \
            const int cs = (p - 2) % J_SLOTS;                                                                         \
            if (ok0) {                                                                                                \
                const int yy = row + 2, zz = zp + 2;                                                                  \
                float JA[3][3], JB[3][3];  /* 12 * (I + grad u) at z and z+1: det(I + J) = det(12 I + 12 J) / 12^3 */  \
                _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                                       \
                    const float2 ym2 = PL2(cs, c, yy - 2, zz), ym1 = PL2(cs, c, yy - 1, zz);                          \
                    const float2 yp1 = PL2(cs, c, yy + 1, zz), yp2 = PL2(cs, c, yy + 2, zz);                          \
                    const float2 zl = PL2(cs, c, yy, zz - 2), zh = PL2(cs, c, yy, zz + 2);                            \
                    const float c0 = win[c][0][(K + 3) % 5], c1 = win[c][1][(K + 3) % 5];                             \
                    JA[c][0] = d4f(win[c][0][(K + 1) % 5], win[c][0][(K + 2) % 5], win[c][0][(K + 4) % 5], win[c][0][K]); \
                    JB[c][0] = d4f(win[c][1][(K + 1) % 5], win[c][1][(K + 2) % 5], win[c][1][(K + 4) % 5], win[c][1][K]); \
                    JA[c][1] = d4f(ym2.x, ym1.x, yp1.x, yp2.x);                                                       \
                    JB[c][1] = d4f(ym2.y, ym1.y, yp1.y, yp2.y);                                                       \
                    JA[c][2] = d4f(zl.x, zl.y, c1, zh.x);                                                             \
                    JB[c][2] = d4f(zl.y, c0, zh.x, zh.y);                                                             \
                }                                                                                                     \
                const double da = det12(JA), db = det12(JB);                                                          \
                if (dp) {                                                                                             \
                    Tout *q = dp + (size_t)(p - 4) * dplane;                                                          \
                    if (pair_store) store_pair(q, da, db);                                                            \
                    else { q[0] = (Tout)da; if (ok1) q[1] = (Tout)db; }                                               \
                }                                                                                                     \
                s += da; s2 += da * da; nn += (!(da > 0.0) && da != 0.0) ? 1.0 : 0.0;                                                 \
                if (ok1) { s += db; s2 += db * db; nn += (!(db > 0.0) && db != 0.0) ? 1.0 : 0.0; }                                    \
            }                                                                                                         \
        }                                                                                                             \
        __syncthreads();   /* every reader of plane p - 2's predecessor slots is done */                              \
        if (threadIdx.x == 0 && p + J_SLOTS - 2 < np) issue(p + J_SLOTS - 2);                                         \
        ++p;                                                                                                          \
    }

    int p = 0;
    while (p < np) {
        JAC_STEP(0) JAC_STEP(1) JAC_STEP(2) JAC_STEP(3) JAC_STEP(4)
    }
#undef JAC_STEP
#undef PL2
#undef PL
    if (!partials) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nn += __shfl_down_sync(0xffffffffu, nn, o);
    }
    __shared__ double sh[3][8];
    if (lane == 0) { sh[0][warp] = nn; sh[1][warp] = s; sh[2][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += sh[threadIdx.x][k];
        const size_t nblk = (size_t)gridDim.x * gridDim.y;
        const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        partials[((size_t)blockIdx.z * 3 + threadIdx.x) * nblk + blk] = a;
    }
}

static size_t jac_tiled_blocks(int X, int Y, int Z) {
    const size_t nzt = (Z - 4 + JT_Z - 1) / JT_Z, nyt = (Y - 4 + JT_Y - 1) / JT_Y, nxt = (X - 4 + JT_X - 1) / JT_X;
    return nzt * nyt * nxt;
}

template <typename Tin, typename Tout>
static int launch_jacdet(const void *field, void *det, double *partials, int B, int X, int Y, int Z,
                         unsigned flags, cudaStream_t st) {
    const uint32_t plane = (uint32_t)(Y - 4) * (Z - 4);
    dim3 grid((plane + 255) / 256, X - 4, B), block(256);
    FastDiv fd = make_fastdiv(Z - 4);
    if (flags & DFM_FIELD_IN_CL)
        k_jacdet<Tin, Tout, true><<<grid, block, 0, st>>>((const Tin *)field, (Tout *)det, partials, X, Y, Z, fd, plane);
    else
        k_jacdet<Tin, Tout, false><<<grid, block, 0, st>>>((const Tin *)field, (Tout *)det, partials, X, Y, Z, fd, plane);
    return check_launch("dfm_jacdet");
}

}  // namespace dfm

using namespace dfm;

extern "C" size_t dfm_jacdet_workspace_bytes(int B, int X, int Y, int Z) {
    if (B <= 0 || X < 5 || Y < 5 || Z < 5) return 0;
    const size_t plane = (size_t)(Y - 4) * (Z - 4);
    size_t nblk = ((plane + 255) / 256) * (size_t)(X - 4);
    if (jac_tiled_blocks(X, Y, Z) > nblk) nblk = jac_tiled_blocks(X, Y, Z);
    return (size_t)B * 3 * nblk * sizeof(double);
}

extern "C" int dfm_jacdet(const void *field, void *det, double *stats, void *partials, int B, int X, int Y,
                          int Z, int in_f64, int out_f64, unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && X >= 5 && Y >= 5 && Z >= 5, DFM_EINVAL,
                "dfm_jacdet: field (%d,%d,%d) needs at least 5 voxels per axis (interior [2:-2])", X, Y, Z);
    DFM_REQUIRE(B <= 65535 && X <= 65535, DFM_EINVAL, "dfm_jacdet: B and X must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 31), DFM_EINVAL, "dfm_jacdet: volume too large");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_jacdet: Y*Z*Z must be < 2^32");
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(field, DFM_EINVAL, "dfm_jacdet: null field");
    DFM_REQUIRE(!stats || partials, DFM_EINVAL, "dfm_jacdet: stats requested without a partials workspace");
    cudaStream_t st = (cudaStream_t)stream;
    double *part = stats ? (double *)partials : nullptr;
    int rc;
    size_t nblk;
    if (!in_f64 && !(flags & DFM_FIELD_IN_CL) && tma_planar_ok((const float *)field, X, Y, Z)) {
        // planar fp32: plane-marching tile kernel
        const int nzt = (Z - 4 + JT_Z - 1) / JT_Z, nyt = (Y - 4 + JT_Y - 1) / JT_Y, nxt = (X - 4 + JT_X - 1) / JT_X;
        dim3 grid(nzt * nyt, nxt, B), block(256);
        CUtensorMap tmap;
        DFM_REQUIRE(encode_planar_map(&tmap, (const float *)field, B * 3, X, Y, Z, 1, JP_Y, JP_Z, 3), DFM_ECUDA,
                    "dfm_jacdet: cuTensorMapEncodeTiled failed");
        if (out_f64) k_jacdet_tiled<double><<<grid, block, 0, st>>>(tmap, (double *)det, part, X, Y, Z, nzt);
        else k_jacdet_tiled<float><<<grid, block, 0, st>>>(tmap, (float *)det, part, X, Y, Z, nzt);
        rc = check_launch("dfm_jacdet(tiled)");
        nblk = jac_tiled_blocks(X, Y, Z);
    } else {
        if (in_f64) rc = out_f64 ? launch_jacdet<double, double>(field, det, part, B, X, Y, Z, flags, st)
                                 : launch_jacdet<double, float>(field, det, part, B, X, Y, Z, flags, st);
        else        rc = out_f64 ? launch_jacdet<float, double>(field, det, part, B, X, Y, Z, flags, st)
                                 : launch_jacdet<float, float>(field, det, part, B, X, Y, Z, flags, st);
        const size_t plane = (size_t)(Y - 4) * (Z - 4);
        nblk = ((plane + 255) / 256) * (size_t)(X - 4);
    }
    if (rc || !stats) return rc;
    k_jacdet_finalize<<<B, 256, 0, st>>>(part, stats, nblk, (double)(X - 4) * (Y - 4) * (Z - 4));
    return check_launch("dfm_jacdet(finalize)");
}
