// Losses adjacent to the warp: the voxel-level sums and gradients of
//   vxm.losses.Dice().loss(map_2, pred)                      [UR]  (train_synthmorph.py:305)
//   vxm.losses.Grad('l2', loss_mult).loss(None, flow)        [UR]  (train_synthmorph.py:306)
// Both are streaming reductions / element-wise kernels bound by HBM; the few scalars per (item, channel)
// are combined on the host side of the ABI (divide_no_nan, means).
//
// Dice: the maps are channels-last [B][N][C] (the reference layout) or planar [B][C][N].  For the
// channels-last case the threads of an item stride through the flat array by a multiple of C, so the
// channel of every element a thread touches is fixed and its partial sums live in registers.
#include "dfm_common.cuh"

namespace dfm {

constexpr int LOSS_THREADS = 256;

// ---------------------------------------------------------------------------------------
// Dice sums.  partial[b][blk][c][2] = (sum t*p, sum t+p) of the elements block `blk` saw.
// ---------------------------------------------------------------------------------------
// VEC = 4: a thread visits float4 groups (16-byte loads); needs total % 4 == 0 and 16-byte aligned maps
template <bool CL, int VEC>
__global__ void __launch_bounds__(LOSS_THREADS)
k_dice_partial(const float *__restrict__ yt, const float *__restrict__ yp, double *__restrict__ partial, int C,
               size_t N, uint32_t stride /* CL: threads per item, a multiple of C */) {
    extern __shared__ float s_acc[];                       // [C][2]
    for (int i = threadIdx.x; i < 2 * C; i += LOSS_THREADS) s_acc[i] = 0.f;
    __syncthreads();
    const size_t total = N * (size_t)C;
    const float *t = yt + (size_t)blockIdx.y * total, *p = yp + (size_t)blockIdx.y * total;
    if (CL) {
        const uint32_t g = blockIdx.x * LOSS_THREADS + threadIdx.x;      // group index of the first visit
        if (g < stride) {
            // the thread's VEC channels never change: the stride (in elements) is a multiple of C
            int c[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) c[j] = (int)(((size_t)g * VEC + j) % (size_t)C);
            float top[VEC], bot[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) top[j] = bot[j] = 0.f;
            const size_t ngroups = total / VEC;
#pragma unroll 4
            for (size_t q = g; q < ngroups; q += stride) {
                float a[VEC], b[VEC];
                if (VEC == 4) {
                    const float4 va = __ldg(reinterpret_cast<const float4 *>(t) + q), vb = __ldg(reinterpret_cast<const float4 *>(p) + q);
                    a[0] = va.x; a[VEC > 1 ? 1 : 0] = va.y; a[VEC > 2 ? 2 : 0] = va.z; a[VEC > 3 ? 3 : 0] = va.w;
                    b[0] = vb.x; b[VEC > 1 ? 1 : 0] = vb.y; b[VEC > 2 ? 2 : 0] = vb.z; b[VEC > 3 ? 3 : 0] = vb.w;
                } else {
                    a[0] = __ldg(t + q); b[0] = __ldg(p + q);
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) { top[j] = fmaf(a[j], b[j], top[j]); bot[j] += a[j] + b[j]; }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) { atomicAdd(&s_acc[2 * c[j]], top[j]); atomicAdd(&s_acc[2 * c[j] + 1], bot[j]); }
        }
    } else {
        // planar: blockIdx.z = channel, the block strides through the channel's N voxels
        const int c = blockIdx.z;
        const float *tc = t + (size_t)c * N, *pc = p + (size_t)c * N;
        float top = 0.f, bot = 0.f;
#pragma unroll 4
        for (size_t e = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; e < N; e += (size_t)gridDim.x * LOSS_THREADS) {
            const float a = __ldg(tc + e), b = __ldg(pc + e);
            top = fmaf(a, b, top);
            bot += a + b;
        }
        atomicAdd(&s_acc[2 * c], top);
        atomicAdd(&s_acc[2 * c + 1], bot);
    }
    __syncthreads();
    // partial[b][i][blk]: the finalize pass reads it coalesced
    const size_t nblk = gridDim.x;
    double *out = partial + (size_t)blockIdx.y * 2 * C * nblk + blockIdx.x;
    if (CL) {
        for (int i = threadIdx.x; i < 2 * C; i += LOSS_THREADS) out[(size_t)i * nblk] = (double)s_acc[i];
    } else if (threadIdx.x < 2) {
        out[(size_t)(2 * blockIdx.z + threadIdx.x) * nblk] = (double)s_acc[2 * blockIdx.z + threadIdx.x];
    }
}

// sums[b][i] = sum over blocks, one CTA per (i, b), fixed summation order (deterministic, fp64)
__global__ void __launch_bounds__(128)
k_dice_finalize(const double *__restrict__ partial, double *__restrict__ sums, int C, int nblk) {
    const double *p = partial + ((size_t)blockIdx.y * 2 * C + blockIdx.x) * nblk;
    double a = 0.0;
    for (int k = threadIdx.x; k < nblk; k += 128) a += p[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    __shared__ double sh[4];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) sums[(size_t)blockIdx.y * 2 * C + blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// d loss / d pred[b, n, c] = alpha[b][c] * true[b, n, c] + beta[b][c]
template <bool CL, int VEC>
__global__ void __launch_bounds__(LOSS_THREADS)
k_dice_bwd(const float *__restrict__ yt, const float *__restrict__ coef, float *__restrict__ g, int C, size_t N) {
    extern __shared__ float s_coef[];                      // [C][2]
    for (int i = threadIdx.x; i < 2 * C; i += LOSS_THREADS) s_coef[i] = __ldg(coef + (size_t)blockIdx.y * 2 * C + i);
    __syncthreads();
    const size_t total = N * (size_t)C;
    const float *t = yt + (size_t)blockIdx.y * total;
    float *o = g + (size_t)blockIdx.y * total;
    const size_t step = (size_t)gridDim.x * LOSS_THREADS;
    if (CL) {
        size_t q = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x;          // group of VEC elements
        int c = (int)((q * VEC) % (size_t)C);               // first channel of the group; advances by a fixed amount
        const int cstep = (int)((step * VEC) % (size_t)C);
        const size_t ngroups = total / VEC;
#pragma unroll 2
        for (; q < ngroups; q += step) {
            int cj[VEC];
            cj[0] = c;
#pragma unroll
            for (int j = 1; j < VEC; ++j) { cj[j] = cj[j - 1] + 1; if (cj[j] >= C) cj[j] -= C; }
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(t) + q);
                float4 r;
                r.x = fmaf(s_coef[2 * cj[0]], v.x, s_coef[2 * cj[0] + 1]);
                r.y = fmaf(s_coef[2 * cj[VEC > 1 ? 1 : 0]], v.y, s_coef[2 * cj[VEC > 1 ? 1 : 0] + 1]);
                r.z = fmaf(s_coef[2 * cj[VEC > 2 ? 2 : 0]], v.z, s_coef[2 * cj[VEC > 2 ? 2 : 0] + 1]);
                r.w = fmaf(s_coef[2 * cj[VEC > 3 ? 3 : 0]], v.w, s_coef[2 * cj[VEC > 3 ? 3 : 0] + 1]);
                reinterpret_cast<float4 *>(o)[q] = r;
            } else {
                o[q] = fmaf(s_coef[2 * c], __ldg(t + q), s_coef[2 * c + 1]);
            }
            c += cstep;
            if (c >= C) c -= C;
        }
    } else {
        const int c = blockIdx.z;
        const float a = s_coef[2 * c], b = s_coef[2 * c + 1];
        const float *tc = t + (size_t)c * N;
        float *oc = o + (size_t)c * N;
#pragma unroll 4
        for (size_t e = (size_t)blockIdx.x * LOSS_THREADS + threadIdx.x; e < N; e += step) oc[e] = fmaf(a, __ldg(tc + e), b);
    }
}

// ---------------------------------------------------------------------------------------
// Grad('l2'): per item and axis the sum of squared forward differences of all three components.
// partial[b][blk][3]; planar or channels-last field.
// ---------------------------------------------------------------------------------------
template <bool CL>
__global__ void __launch_bounds__(LOSS_THREADS)
k_grad_partial(const float *__restrict__ f, double *__restrict__ partial, int X, int Y, int Z, FastDiv zdiv, FastDiv ydiv) {
    const uint32_t N = (uint32_t)X * Y * Z;
    const float *fb = f + (size_t)blockIdx.y * 3 * N;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (uint32_t n = blockIdx.x * LOSS_THREADS + threadIdx.x; n < N; n += gridDim.x * LOSS_THREADS) {
        const uint32_t q = fast_div(n, zdiv), z = n - q * zdiv.d, x = fast_div(q, ydiv), y = q - x * ydiv.d;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *pc = CL ? fb + c : fb + (size_t)c * N;
            const uint32_t es = CL ? 3u : 1u;
            const float v = __ldg(pc + (size_t)n * es);
            if (x + 1 < (uint32_t)X) { const float d = __ldg(pc + (size_t)(n + (uint32_t)Y * Z) * es) - v; sx = fmaf(d, d, sx); }
            if (y + 1 < (uint32_t)Y) { const float d = __ldg(pc + (size_t)(n + (uint32_t)Z) * es) - v; sy = fmaf(d, d, sy); }
            if (z + 1 < (uint32_t)Z) { const float d = __ldg(pc + (size_t)(n + 1u) * es) - v; sz = fmaf(d, d, sz); }
        }
    }
    __shared__ double sh[3][LOSS_THREADS / 32];
    double a[3] = {(double)sx, (double)sy, (double)sz};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = a[0]; sh[1][warp] = a[1]; sh[2][warp] = a[2]; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < LOSS_THREADS / 32; ++k) t += sh[threadIdx.x][k];
        partial[((size_t)blockIdx.y * 3 + threadIdx.x) * gridDim.x + blockIdx.x] = t;      // [b][axis][blk]
    }
}

__global__ void __launch_bounds__(128)
k_grad_finalize(const double *__restrict__ partial, double *__restrict__ sums, int nblk) {
    const double *p = partial + ((size_t)blockIdx.y * 3 + blockIdx.x) * nblk;
    double a = 0.0;
    for (int k = threadIdx.x; k < nblk; k += 128) a += p[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    __shared__ double sh[4];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) sums[(size_t)blockIdx.y * 3 + blockIdx.x] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// g[b, n, c] = sum_axis coef[b][axis] * ((v[n] - v[n - e]) - (v[n + e] - v[n]))   (missing neighbours drop out)
template <bool CL>
__global__ void __launch_bounds__(LOSS_THREADS)
k_grad_bwd(const float *__restrict__ f, const float *__restrict__ coef, float *__restrict__ g, int X, int Y, int Z,
           FastDiv zdiv, FastDiv ydiv) {
    const uint32_t N = (uint32_t)X * Y * Z;
    const float *fb = f + (size_t)blockIdx.y * 3 * N;
    float *gb = g + (size_t)blockIdx.y * 3 * N;
    const float cx = __ldg(coef + blockIdx.y * 3), cy = __ldg(coef + blockIdx.y * 3 + 1), cz = __ldg(coef + blockIdx.y * 3 + 2);
    const uint32_t SX = (uint32_t)Y * Z, SY = (uint32_t)Z;
    for (uint32_t n = blockIdx.x * LOSS_THREADS + threadIdx.x; n < N; n += gridDim.x * LOSS_THREADS) {
        const uint32_t q = fast_div(n, zdiv), z = n - q * zdiv.d, x = fast_div(q, ydiv), y = q - x * ydiv.d;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *pc = CL ? fb + c : fb + (size_t)c * N;
            const uint32_t es = CL ? 3u : 1u;
            const float v = __ldg(pc + (size_t)n * es);
            float a = 0.f;
            if (x > 0) a = fmaf(cx, v - __ldg(pc + (size_t)(n - SX) * es), a);
            if (x + 1 < (uint32_t)X) a = fmaf(cx, v - __ldg(pc + (size_t)(n + SX) * es), a);
            if (y > 0) a = fmaf(cy, v - __ldg(pc + (size_t)(n - SY) * es), a);
            if (y + 1 < (uint32_t)Y) a = fmaf(cy, v - __ldg(pc + (size_t)(n + SY) * es), a);
            if (z > 0) a = fmaf(cz, v - __ldg(pc + (size_t)(n - 1u) * es), a);
            if (z + 1 < (uint32_t)Z) a = fmaf(cz, v - __ldg(pc + (size_t)(n + 1u) * es), a);
            (CL ? gb + c : gb + (size_t)c * N)[(size_t)n * es] = a;
        }
    }
}

// blocks per item of the reductions (a few waves of the 148 SMs; every block writes one partial row)
static int loss_blocks(size_t work_items) {
    const size_t want = (work_items + LOSS_THREADS * 16 - 1) / (LOSS_THREADS * 16);
    return (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}
// channels-last Dice: threads per item = a multiple of C that fills a few waves of the SMs
static uint32_t dice_stride(int C, size_t ngroups) {
    uint32_t m = (uint32_t)((148u * 4u * LOSS_THREADS) / (uint32_t)C);
    if (m < 1) m = 1;
    uint64_t s = (uint64_t)m * (uint32_t)C;
    if (s > ngroups) s = ((ngroups + C - 1) / C) * (uint64_t)C;
    return (uint32_t)s;
}
int launch_dice_finalize(const double *partial, double *sums, int B, int C, int nblk, cudaStream_t st) {
    dim3 fgrid(2 * C, B);
    k_dice_finalize<<<fgrid, 128, 0, st>>>(partial, sums, C, nblk);
    return check_launch("dice finalize");
}
static bool vec4_ok(const void *a, const void *b, size_t total) {
    return total % 4 == 0 && (((uintptr_t)a | (uintptr_t)b) & 15u) == 0;
}

}  // namespace dfm

using namespace dfm;

extern "C" size_t dfm_dice_workspace_bytes(int B, int C, size_t N) {
    if (B <= 0 || C <= 0 || N == 0) return 0;
    const uint32_t stride = dice_stride(C, N * (size_t)C);          // upper bound over both vector widths
    const size_t cl_blocks = (stride + LOSS_THREADS - 1) / LOSS_THREADS;
    const size_t pl_blocks = (size_t)loss_blocks(N);
    const size_t nblk = cl_blocks > pl_blocks ? cl_blocks : pl_blocks;
    return (size_t)B * nblk * 2 * C * sizeof(double);
}

extern "C" int dfm_dice_sums(const float *y_true, const float *y_pred, double *sums, void *work, int B, int C, size_t N,
                             unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && C <= 2048 && B <= 65535, DFM_EINVAL, "dfm_dice_sums: bad shape B=%d C=%d", B, C);
    DFM_REQUIRE(N * (size_t)C < (1ull << 40), DFM_EINVAL, "dfm_dice_sums: volume too large");
    if (B == 0 || N == 0) return DFM_OK;
    DFM_REQUIRE(y_true && y_pred && sums && work, DFM_EINVAL, "dfm_dice_sums: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = (double *)work;
    const size_t smem = 2 * (size_t)C * sizeof(float);
    const size_t total = N * (size_t)C;
    int nblk;
    if ((flags & DFM_IMG_CL) && C > 1) {
        const bool v4 = vec4_ok(y_true, y_pred, total);
        const uint32_t stride = dice_stride(C, v4 ? total / 4 : total);
        nblk = (int)((stride + LOSS_THREADS - 1) / LOSS_THREADS);
        dim3 grid(nblk, B);
        if (v4) k_dice_partial<true, 4><<<grid, LOSS_THREADS, smem, st>>>(y_true, y_pred, partial, C, N, stride);
        else k_dice_partial<true, 1><<<grid, LOSS_THREADS, smem, st>>>(y_true, y_pred, partial, C, N, stride);
    } else {
        DFM_REQUIRE(C <= 65535, DFM_EINVAL, "dfm_dice_sums: C too large");
        nblk = loss_blocks(N);
        dim3 grid(nblk, B, C);
        k_dice_partial<false, 1><<<grid, LOSS_THREADS, smem, st>>>(y_true, y_pred, partial, C, N, 0u);
    }
    int rc = check_launch("dfm_dice_sums");
    if (rc) return rc;
    return launch_dice_finalize(partial, sums, B, C, nblk, st);
}

extern "C" int dfm_dice_bwd(const float *y_true, const float *coef, float *g_pred, int B, int C, size_t N, unsigned flags,
                            void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && C <= 2048 && B <= 65535, DFM_EINVAL, "dfm_dice_bwd: bad shape B=%d C=%d", B, C);
    if (B == 0 || N == 0) return DFM_OK;
    DFM_REQUIRE(y_true && coef && g_pred, DFM_EINVAL, "dfm_dice_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = 2 * (size_t)C * sizeof(float);
    const size_t total = N * (size_t)C;
    if ((flags & DFM_IMG_CL) && C > 1) {
        if (vec4_ok(y_true, g_pred, total)) {
            dim3 grid(loss_blocks(total / 4), B);
            k_dice_bwd<true, 4><<<grid, LOSS_THREADS, smem, st>>>(y_true, coef, g_pred, C, N);
        } else {
            dim3 grid(loss_blocks(total), B);
            k_dice_bwd<true, 1><<<grid, LOSS_THREADS, smem, st>>>(y_true, coef, g_pred, C, N);
        }
    } else {
        DFM_REQUIRE(C <= 65535, DFM_EINVAL, "dfm_dice_bwd: C too large");
        dim3 grid(loss_blocks(N), B, C);
        k_dice_bwd<false, 1><<<grid, LOSS_THREADS, smem, st>>>(y_true, coef, g_pred, C, N);
    }
    return check_launch("dfm_dice_bwd");
}

extern "C" size_t dfm_grad_l2_workspace_bytes(int B, int X, int Y, int Z) {
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0) return 0;
    return (size_t)B * loss_blocks((size_t)X * Y * Z) * 3 * sizeof(double);
}

extern "C" int dfm_grad_l2_sums(const float *flow, double *sums, void *work, int B, int X, int Y, int Z, unsigned flags,
                                void *stream) {
    int rc;
    if ((rc = validate_grid("dfm_grad_l2_sums", B, X, Y, Z))) return rc;
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(flow && sums && work, DFM_EINVAL, "dfm_grad_l2_sums: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = loss_blocks((size_t)X * Y * Z);
    dim3 grid(nblk, B);
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    if (flags & DFM_FIELD_IN_CL) k_grad_partial<true><<<grid, LOSS_THREADS, 0, st>>>(flow, (double *)work, X, Y, Z, zd, yd);
    else k_grad_partial<false><<<grid, LOSS_THREADS, 0, st>>>(flow, (double *)work, X, Y, Z, zd, yd);
    if ((rc = check_launch("dfm_grad_l2_sums"))) return rc;
    k_grad_finalize<<<dim3(3, B), 128, 0, st>>>((const double *)work, sums, nblk);
    return check_launch("dfm_grad_l2_sums(finalize)");
}

extern "C" int dfm_grad_l2_bwd(const float *flow, const float *coef, float *g, int B, int X, int Y, int Z, unsigned flags,
                               void *stream) {
    int rc;
    if ((rc = validate_grid("dfm_grad_l2_bwd", B, X, Y, Z))) return rc;
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(flow && coef && g && g != flow, DFM_EINVAL, "dfm_grad_l2_bwd: null or aliased pointer");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(loss_blocks((size_t)X * Y * Z), B);
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    if (flags & DFM_FIELD_IN_CL) k_grad_bwd<true><<<grid, LOSS_THREADS, 0, st>>>(flow, coef, g, X, Y, Z, zd, yd);
    else k_grad_bwd<false><<<grid, LOSS_THREADS, 0, st>>>(flow, coef, g, X, Y, Z, zd, yd);
    return check_launch("dfm_grad_l2_bwd");
}

// ---------------------------------------------------------------------------------------
// d Dice / d field through SpatialTransformer('linear') without materialising d Dice / d pred
// (dfm_warp_cl.cu kernel with the upstream gradient formed on the fly)
// ---------------------------------------------------------------------------------------
extern "C" int dfm_warp_dice_bwd(const float *y_true, const float *coef, const float *img, const float *field,
                                 float *gfield, int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill,
                                 unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && X >= 1 && Y >= 1 && Z >= 1 && B <= 65535, DFM_EINVAL,
                "dfm_warp_dice_bwd: bad shape");
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(y_true && coef && img && field && gfield, DFM_EINVAL, "dfm_warp_dice_bwd: null pointer");
    int rc = launch_warp_cl_dice_bwd(y_true, coef, img, field, gfield, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, flags, (cudaStream_t)stream);
    if (rc == DFM_EUNSUPPORTED) return fail(DFM_EUNSUPPORTED, "dfm_warp_dice_bwd: C = %d or shape not supported by the fused kernel", C);
    return rc;
}
