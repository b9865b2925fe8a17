// Jacobian-determinant map of a displacement field (eval_reg_with_jacobian.py:62-78):
// 4th-order central differences on the interior [2:-2]^3, det(I + J) in fp64, fold count and
// moments.  fp64 in registers is free under the memory roof; HBM sees the field once
// (12 B/voxel as fp32) and the determinant map once.
#include "dfm_common.cuh"

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
    double nn = (valid && dval < 0.0) ? 1.0 : 0.0;
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
    const size_t nblk = ((plane + 255) / 256) * (size_t)(X - 4);
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
    if (in_f64) rc = out_f64 ? launch_jacdet<double, double>(field, det, part, B, X, Y, Z, flags, st)
                             : launch_jacdet<double, float>(field, det, part, B, X, Y, Z, flags, st);
    else        rc = out_f64 ? launch_jacdet<float, double>(field, det, part, B, X, Y, Z, flags, st)
                             : launch_jacdet<float, float>(field, det, part, B, X, Y, Z, flags, st);
    if (rc || !stats) return rc;
    const size_t plane = (size_t)(Y - 4) * (Z - 4);
    const size_t nblk = ((plane + 255) / 256) * (size_t)(X - 4);
    k_jacdet_finalize<<<B, 256, 0, st>>>(part, stats, nblk, (double)(X - 4) * (Y - 4) * (Z - 4));
    return check_launch("dfm_jacdet(finalize)");
}
