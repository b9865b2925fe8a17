// Linear warp of a ONE-HOT label map, forward and d/d field, from the label map itself:
//   pred = SpatialTransformer('linear')([one_hot(labels, C), flow])            (train_synthmorph.py:298; map_1 is the one-hot
//   output of ne.models.labels_to_image, :288-289, so it is one-hot by construction)
// The generic channels-last kernel (dfm_warp_cl.cu) gathers 8 corners x C floats per voxel (832 B at C = 26) through the L1 to
// multiply almost all of them by zero.  A one-hot corner IS its label, so here a voxel gathers 8 BYTES:
//   forward   pred[n, c] = sum of the corner weights w_k whose label is c, accumulated in corner order -- exactly what
//             tri_accumulate computes on 0/1 values (w * 0 adds nothing, w * 1 is w), so the result is BIT-IDENTICAL to the
//             generic kernel on the materialised one-hot tensor, in both builds; the 2 x 511 MB one-hot tensor of a
//             160 x 160 x 192 x 26 item is never built or read;
//   backward  sum_c g[n, c] * v_k[c] is g[n, label_k]: the upstream gradient tile of 256 voxels is staged in shared memory
//             (one coalesced read), each voxel picks its 8 values and combines them with the weight derivatives (TensorFlow
//             autodiff semantics as in k_warp_cl_bwd: floor has no gradient, clip passes it on [0, max]).  No atomics
//             (there is no gradient with respect to labels).
// Phase A (thread = voxel): location, weights, the 8 corner labels (uint8 gathers, L1 hits between neighbours) -> a 48-byte
// record in shared memory.  Phase B (thread = output element of the block's contiguous 256 x C tile): coalesced stores.
#include "dfm_common.cuh"

namespace dfm {

constexpr int ONEHOT_BWD_CMAX = 48;      // largest C whose [256][C] float tile fits 48 KB of shared memory

struct __align__(16) OneHotRec {
    float w[8];
    uint32_t lab[2];    // corner labels 0-3 / 4-7, one byte each (corner order)
    uint32_t dead;      // fill_value applies
    uint32_t pad;
};

__device__ __forceinline__ void onehot_loc(const float *__restrict__ fb, uint32_t n, uint32_t N, bool field_cl, int Y, int Z,
                                           float &lx, float &ly, float &lz) {
    if (field_cl) { lx = __ldg(fb + (size_t)n * 3); ly = __ldg(fb + (size_t)n * 3 + 1); lz = __ldg(fb + (size_t)n * 3 + 2); }
    else { lx = __ldg(fb + n); ly = __ldg(fb + N + n); lz = __ldg(fb + 2 * (size_t)N + n); }
    const uint32_t q = n / (uint32_t)Z, z = n - q * (uint32_t)Z, x = q / (uint32_t)Y, y = q - x * (uint32_t)Y;
    lx = __fadd_rn((float)x, lx); ly = __fadd_rn((float)y, ly); lz = __fadd_rn((float)z, lz);
}

__device__ __forceinline__ void onehot_labels(const uint8_t *__restrict__ lb, uint32_t base, uint32_t oZ, uint32_t oY, uint32_t (&lab)[2]) {
    const uint8_t *p = lb + base;
    lab[0] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + oZ) << 16) | ((uint32_t)__ldg(p + oZ + 1) << 24);
    p += oY;
    lab[1] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + oZ) << 16) | ((uint32_t)__ldg(p + oZ + 1) << 24);
}

template <bool HF>
__global__ void __launch_bounds__(256)
k_warp_onehot(const uint8_t *__restrict__ labels, const float *__restrict__ field, float *__restrict__ out, int C, int Xi, int Yi,
              int Zi, int Y, int Z, uint32_t N, float fill, int field_cl, FastDiv cdiv) {
    __shared__ OneHotRec s_rec[256];
    const uint32_t n0 = blockIdx.x * 256u;
    {   // ---- phase A: thread = voxel -------------------------------------------------------
        const uint32_t n = min(n0 + threadIdx.x, N - 1);
        float lx, ly, lz;
        onehot_loc(field + (size_t)blockIdx.y * 3 * N, n, N, field_cl, Y, Z, lx, ly, lz);
        const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
        const AxisF ax = axis_fast(lx, (float)mxi, mxi), ay = axis_fast(ly, (float)myi, myi), az = axis_fast(lz, (float)mzi, mzi);
        OneHotRec r;
        tri_weights(ax, ay, az, r.w);
        const uint32_t base = ((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1);
        onehot_labels(labels + (size_t)blockIdx.y * ((size_t)Xi * Yi * Zi), base, (uint32_t)Zi, (uint32_t)Yi * Zi, r.lab);
        r.dead = HF && (lx < 0.f || lx > (float)mxi || ly < 0.f || ly > (float)myi || lz < 0.f || lz > (float)mzi);
        r.pad = 0;
        s_rec[threadIdx.x] = r;
    }
    __syncthreads();
    // ---- phase B: thread = element of the block's contiguous [256][C] output tile ------------
    const uint32_t nv = min(256u, N - n0), ne = nv * (uint32_t)C;
    float *po = out + ((size_t)blockIdx.y * N + n0) * C;
    for (uint32_t e = threadIdx.x; e < ne; e += 256u) {
        const uint32_t v = fast_div(e, cdiv), c = e - v * (uint32_t)C;
        const float4 wa = *reinterpret_cast<const float4 *>(&s_rec[v].w[0]);
        const float4 wb = *reinterpret_cast<const float4 *>(&s_rec[v].w[4]);
        const uint4 t = *reinterpret_cast<const uint4 *>(&s_rec[v].lab[0]);
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        // tri_accumulate on 0/1 values: acc = w0 * v0, then acc = w_k * v_k + acc (one rounding per step in both builds)
        float acc = ((t.x & 0xffu) == c) ? w[0] : 0.f;
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const uint32_t l = ((k < 4 ? t.x : t.y) >> (8 * (k & 3))) & 0xffu;
            if (l == c) acc = __fadd_rn(acc, w[k]);
        }
        if (HF && t.z) acc = fill;
        __stcs(po + e, acc);
    }
}

// The same forward pass for C <= 48 through a shared [256][C] output tile: a one-hot row has at most 8 non-zero entries, so
// instead of evaluating 8 compares for each of the C outputs of a voxel (~35 instructions per output element: the kernel above
// is issue-bound at 0.51 ms for 2 x 160 x 160 x 192 x 26), the tile is zeroed, every thread adds its voxel's 8 corner weights
// into its own row IN CORNER ORDER (0 + w_k is exact, later additions round once as in tri_accumulate: same bits), and the
// tile leaves with coalesced 16-byte stores.
template <bool HF>
__global__ void __launch_bounds__(256)
k_warp_onehot_tile(const uint8_t *__restrict__ labels, const float *__restrict__ field, float *__restrict__ out, int C, int Xi,
                   int Yi, int Zi, int Y, int Z, uint32_t N, float fill, int field_cl) {
    extern __shared__ __align__(16) float s_out[];            // [256][C]
    const uint32_t n0 = blockIdx.x * 256u;
    const uint32_t nv = min(256u, N - n0), ne = nv * (uint32_t)C;
    for (uint32_t e = threadIdx.x * 4u; e < 256u * (uint32_t)C; e += 1024u)          // 256 * C is a multiple of 4
        *reinterpret_cast<float4 *>(s_out + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t n = min(n0 + threadIdx.x, N - 1);
    float lx, ly, lz;
    onehot_loc(field + (size_t)blockIdx.y * 3 * N, n, N, field_cl, Y, Z, lx, ly, lz);
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const AxisF ax = axis_fast(lx, (float)mxi, mxi), ay = axis_fast(ly, (float)myi, myi), az = axis_fast(lz, (float)mzi, mzi);
    float w[8];
    tri_weights(ax, ay, az, w);
    uint32_t lab[2];
    const uint32_t base = ((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1);
    onehot_labels(labels + (size_t)blockIdx.y * ((size_t)Xi * Yi * Zi), base, (uint32_t)Zi, (uint32_t)Yi * Zi, lab);
    const bool dead = HF && (lx < 0.f || lx > (float)mxi || ly < 0.f || ly > (float)myi || lz < 0.f || lz > (float)mzi);
    __syncthreads();
    float *row = s_out + threadIdx.x * (uint32_t)C;           // this thread's row: no other thread touches it before the barrier
    if (HF && dead) {
        for (int c = 0; c < C; ++c) row[c] = fill;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t l = ((k < 4 ? lab[0] : lab[1]) >> (8 * (k & 3))) & 0xffu;
            if (l < (uint32_t)C) row[l] = __fadd_rn(row[l], w[k]);
        }
    }
    __syncthreads();
    float *po = out + ((size_t)blockIdx.y * N + n0) * C;
    if (((reinterpret_cast<uintptr_t>(po) | (uintptr_t)(ne * 4u)) & 15u) == 0) {
        for (uint32_t e = threadIdx.x * 4u; e < ne; e += 1024u) __stcs(reinterpret_cast<float4 *>(po + e), *reinterpret_cast<const float4 *>(s_out + e));
    } else {
        for (uint32_t e = threadIdx.x; e < ne; e += 256u) __stcs(po + e, s_out[e]);
    }
}

// d/d field.  CMAX bounds the shared gradient tile (256 voxels x C floats)

__global__ void __launch_bounds__(256)
k_warp_onehot_bwd(const float *__restrict__ gout, const uint8_t *__restrict__ labels, const float *__restrict__ field,
                  float *__restrict__ gfield, int C, int Xi, int Yi, int Zi, int Y, int Z, uint32_t N, int has_fill, int field_cl,
                  int gfield_cl) {
    extern __shared__ __align__(16) float s_g[];              // [256][C] upstream gradient of the block's voxels
    const uint32_t n0 = blockIdx.x * 256u;
    const uint32_t nv = min(256u, N - n0), ne = nv * (uint32_t)C;
    const float *pg = gout + ((size_t)blockIdx.y * N + n0) * C;
    for (uint32_t e = threadIdx.x; e < ne; e += 256u) s_g[e] = __ldcs(pg + e);
    __syncthreads();
    const uint32_t n = n0 + threadIdx.x;
    if (n >= N) return;
    float lx, ly, lz;
    onehot_loc(field + (size_t)blockIdx.y * 3 * N, n, N, field_cl, Y, Z, lx, ly, lz);
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const AxisF ax = axis_fast(lx, mxf, mxi), ay = axis_fast(ly, myf, myi), az = axis_fast(lz, mzf, mzi);
    const bool dead = has_fill && (lx < 0.f || lx > mxf || ly < 0.f || ly > myf || lz < 0.f || lz > mzf);
    // clip passes the gradient on [0, max]; at loc == max both reference corners are the edge voxel (difference 0),
    // which the i1 - 1 addressing reproduces with a strict upper bound (as k_warp_cl_bwd)
    const float gx = (!dead && lx >= 0.f && lx < mxf) ? 1.f : 0.f;
    const float gy = (!dead && ly >= 0.f && ly < myf) ? 1.f : 0.f;
    const float gz = (!dead && lz >= 0.f && lz < mzf) ? 1.f : 0.f;
    uint32_t lab[2];
    const uint32_t base = ((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1);
    onehot_labels(labels + (size_t)blockIdx.y * ((size_t)Xi * Yi * Zi), base, (uint32_t)Zi, (uint32_t)Yi * Zi, lab);
    const float *row = s_g + threadIdx.x * (uint32_t)C;
    float s[8];                                               // s_k = sum_c g[c] * onehot(label_k)[c] = g[label_k]
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t l = ((k < 4 ? lab[0] : lab[1]) >> (8 * (k & 3))) & 0xffu;
        s[k] = l < (uint32_t)C ? row[l] : 0.f;                // a label outside [0, C) is an all-zero one-hot row
    }
    const float wx[2] = {ax.w0, ax.w1}, wy[2] = {ay.w0, ay.w1}, wz[2] = {az.w0, az.w1};
    // corner k = a * 4 + b * 2 + q (x, y, z bits)
    float dx = 0.f, dy = 0.f, dz = 0.f;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            dx = fmaf(wy[p] * wz[q], s[4 + p * 2 + q] - s[p * 2 + q], dx);
            dy = fmaf(wx[p] * wz[q], s[p * 4 + 2 + q] - s[p * 4 + q], dy);
            dz = fmaf(wx[p] * wy[q], s[p * 4 + q * 2 + 1] - s[p * 4 + q * 2], dz);
        }
    dx *= gx; dy *= gy; dz *= gz;
    float *gf = gfield + (size_t)blockIdx.y * 3 * N;
    if (gfield_cl) { gf[(size_t)n * 3] = dx; gf[(size_t)n * 3 + 1] = dy; gf[(size_t)n * 3 + 2] = dz; }
    else { gf[n] = dx; gf[N + n] = dy; gf[2 * (size_t)N + n] = dz; }
}

}  // namespace dfm

using namespace dfm;

static int onehot_check(const char *who, int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 2 && Yi >= 2 && Zi >= 2 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL,
                "%s: bad shape B=%d C=%d labels=(%d,%d,%d) grid=(%d,%d,%d) (the label volume needs every axis >= 2)", who, B, C, Xi, Yi,
                Zi, X, Y, Z);
    DFM_REQUIRE(C <= 256, DFM_EINVAL, "%s: C = %d labels do not fit a byte", who, C);
    DFM_REQUIRE(B <= 65535, DFM_EINVAL, "%s: B must be <= 65535", who);
    DFM_REQUIRE((uint64_t)X * Y * Z < (1ull << 30) && (uint64_t)Xi * Yi * Zi < (1ull << 30), DFM_EINVAL,
                "%s: volume too large (>= 2^30 voxels)", who);
    DFM_REQUIRE((uint64_t)256 * C * (uint64_t)C < (1ull << 32), DFM_EINVAL, "%s: C too large", who);
    return DFM_OK;
}

extern "C" int dfm_warp_onehot_fwd(const uint8_t *labels, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi,
                                   int X, int Y, int Z, int has_fill, float fill, unsigned flags, void *stream) {
    int rc = onehot_check("dfm_warp_onehot_fwd", B, C, Xi, Yi, Zi, X, Y, Z);
    if (rc) return rc;
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(labels && field && out, DFM_EINVAL, "dfm_warp_onehot_fwd: null pointer");
    const uint32_t N = (uint32_t)X * Y * Z;
    dim3 grid((N + 255) / 256, B), block(256);
    const int fcl = (flags & DFM_FIELD_IN_CL) ? 1 : 0;
    const FastDiv cd = make_fastdiv((uint32_t)C);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool no_tile = getenv("DFM_ONEHOT_NO_TILE") != nullptr;                          // tuning aid
    if (C <= ONEHOT_BWD_CMAX && !no_tile) {                   // [256][C] tile <= 48 KB
        const size_t smem = (size_t)256 * C * sizeof(float);
        if (has_fill) k_warp_onehot_tile<true><<<grid, block, smem, st>>>(labels, field, out, C, Xi, Yi, Zi, Y, Z, N, fill, fcl);
        else k_warp_onehot_tile<false><<<grid, block, smem, st>>>(labels, field, out, C, Xi, Yi, Zi, Y, Z, N, fill, fcl);
        return check_launch("k_warp_onehot_tile");
    }
    if (has_fill) k_warp_onehot<true><<<grid, block, 0, st>>>(labels, field, out, C, Xi, Yi, Zi, Y, Z, N, fill, fcl, cd);
    else k_warp_onehot<false><<<grid, block, 0, st>>>(labels, field, out, C, Xi, Yi, Zi, Y, Z, N, fill, fcl, cd);
    return check_launch("k_warp_onehot");
}

extern "C" int dfm_warp_onehot_bwd(const float *gout, const uint8_t *labels, const float *field, float *gfield, int B, int C,
                                   int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags, void *stream) {
    int rc = onehot_check("dfm_warp_onehot_bwd", B, C, Xi, Yi, Zi, X, Y, Z);
    if (rc) return rc;
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(gout && labels && field && gfield, DFM_EINVAL, "dfm_warp_onehot_bwd: null pointer");
    DFM_REQUIRE(C <= ONEHOT_BWD_CMAX, DFM_EUNSUPPORTED, "dfm_warp_onehot_bwd: C = %d > %d (shared gradient tile)", C, ONEHOT_BWD_CMAX);
    const uint32_t N = (uint32_t)X * Y * Z;
    dim3 grid((N + 255) / 256, B), block(256);
    const size_t smem = (size_t)256 * C * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_onehot_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)256 * ONEHOT_BWD_CMAX * sizeof(float)));
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_onehot_bwd smem attribute: %s", cudaGetErrorString(e));
        configured = (size_t)256 * ONEHOT_BWD_CMAX * sizeof(float);
    }
    k_warp_onehot_bwd<<<grid, block, smem, (cudaStream_t)stream>>>(gout, labels, field, gfield, C, Xi, Yi, Zi, Y, Z, N, has_fill,
                                                                  (flags & DFM_FIELD_IN_CL) ? 1 : 0, (flags & DFM_FIELD_OUT_CL) ? 1 : 0);
    return check_launch("k_warp_onehot_bwd");
}
