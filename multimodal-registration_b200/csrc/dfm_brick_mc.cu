// Multi-channel linear image warp through a TMA channel ring (planar fp32):
//   forward   out[b,c,p]   = interp(img[b,c], p + field[b,:,p])                      (C channels)
//   backward  gfield[b,d,p] = sum_c gout[b,c,p] * d interp(img[b,c], .) / d loc_d    (gather only)
// This is the `pred = SpatialTransformer('linear')([one_hot_map, flow])` of the training step
// (train_synthmorph.py:298, C = 26) and its gradient w.r.t. the flow (:305-306), the largest
// single op of the hot path (220 / 232 B per voxel).
//
// The corner indices and weights depend on the field only, so they are computed once per voxel
// and kept in registers; the channels then stream through a ring of shared-memory bricks, one
// 4-D TMA box per channel at the SAME bounding-box origin (prefetch distance NSLOT-1, one
// mbarrier per slot, one __syncthreads per channel).  Per voxel and channel the kernel issues 8
// LDS with immediate offsets + 8 FMA: the shared-memory crossbar (9 wavefronts per 32 voxels and
// channel) stays below the HBM time of the same data (220 B/voxel), which the direct-gather
// kernel (8 L1 requests of 2-3 lines each) cannot do.
// Backward uses  d out / d loc_d = sum_k (d w_k / d loc_d) * (sum_c g_c * v_ck): the channel
// reduction is done first (8 accumulators), so the work per channel equals the forward's.
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

namespace dfm {

constexpr int MT_X = 4, MT_Y = 8, MT_Z = 32;

template <int BX, int BY, int BZ, int NSLOT, bool BWD>
__global__ void __launch_bounds__(256)
k_warp_mc_brick(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ img,
                const float *__restrict__ field, float *__restrict__ out, const float *__restrict__ gout,
                float *__restrict__ gfield, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill,
                float fill, int nzt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PX = BY * BZ, PY = BZ, CS = BX * BY * BZ;       // CS * 4 bytes is a multiple of 128
    float *ring = reinterpret_cast<float *>(smem_raw);            // [NSLOT][BX][BY][BZ]
    __shared__ __align__(8) uint64_t bar[NSLOT];
    __shared__ int s_min[3], s_max[3];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int z = zt * MT_Z + lane, y = yt * MT_Y + warp, x0 = blockIdx.y * MT_X;
    const bool ok_yz = (z < Z) && (y < Y);
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fy = (float)y, fz = (float)z, fx0 = (float)x0;
    const uint32_t vox0 = ((uint32_t)x0 * Y + y) * Z + z, XS = (uint32_t)Y * Z;
    const int nx = ok_yz ? min(MT_X, X - x0) : 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NSLOT; ++k) mbar_init(&bar[k], 1);
        s_min[0] = s_min[1] = s_min[2] = INT_MAX;
        s_max[0] = s_max[1] = s_max[2] = INT_MIN;
    }
    __syncthreads();

    // ---- pass 1: sample locations, bounding box -------------------------------------------------
    float l[3][MT_X];
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
    for (int i = 0; i < MT_X; ++i) {
        l[0][i] = l[1][i] = l[2][i] = 0.f;
        if (i < nx) {
            const uint32_t vox = vox0 + i * XS;
            l[0][i] = __fadd_rn(fx0 + (float)i, __ldg(fb + vox));
            l[1][i] = __fadd_rn(fy, __ldg(fb + N + vox));
            l[2][i] = __fadd_rn(fz, __ldg(fb + 2 * (size_t)N + vox));
            const int ix = axis_fast_i1(l[0][i], mxf, mxi), iy = axis_fast_i1(l[1][i], myf, myi), iz = axis_fast_i1(l[2][i], mzf, mzi);
            mn[0] = min(mn[0], ix); mn[1] = min(mn[1], iy); mn[2] = min(mn[2], iz);
            mx[0] = max(mx[0], ix); mx[1] = max(mx[1], iy); mx[2] = max(mx[2], iz);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        mn[d] = __reduce_min_sync(0xffffffffu, mn[d]);
        mx[d] = __reduce_max_sync(0xffffffffu, mx[d]);
    }
    if (lane == 0 && mn[0] != INT_MAX) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { atomicMin(&s_min[d], mn[d] - 1); atomicMax(&s_max[d], mx[d]); }
    }
    __syncthreads();
    const int ox = s_min[0], oy = s_min[1], oz = s_min[2] & ~3;
    if (ox == INT_MAX) return;
    const bool fits = (s_max[0] - ox < BX) && (s_max[1] - oy < BY) && (s_max[2] - oz < BZ);   // uniform

    // per-voxel set-up kept across the channel loop
    int base[MT_X];
    float w[MT_X][8];
    AxisF ax[MT_X], ay[MT_X], az[MT_X];
    bool dead[MT_X];
#pragma unroll
    for (int i = 0; i < MT_X; ++i) {
        ax[i] = axis_fast(l[0][i], mxf, mxi); ay[i] = axis_fast(l[1][i], myf, myi); az[i] = axis_fast(l[2][i], mzf, mzi);
        tri_weights(ax[i], ay[i], az[i], w[i]);
        dead[i] = has_fill && (l[0][i] < 0.f || l[0][i] > mxf || l[1][i] < 0.f || l[1][i] > myf || l[2][i] < 0.f || l[2][i] > mzf);
        base[i] = fits ? (ax[i].i1 * PX + ay[i].i1 * PY + az[i].i1 - ((ox + 1) * PX + (oy + 1) * PY + (oz + 1)))
                       : (int)(((uint32_t)(ax[i].i1 - 1) * Yi + (uint32_t)(ay[i].i1 - 1)) * Zi + (uint32_t)(az[i].i1 - 1));
    }
    float acc[MT_X][8];                               // backward: sum_c g_c * v_ck
    if (BWD) {
#pragma unroll
        for (int i = 0; i < MT_X; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
    }

    const int vol0 = (int)blockIdx.z * C;
    auto issue = [&](int c) {                         // thread 0 only
        mbar_expect_tx(&bar[c % NSLOT], (uint32_t)(CS * sizeof(float)));
        tma_load_4d(ring + (size_t)(c % NSLOT) * CS, &tmap, &bar[c % NSLOT], oz, oy, ox, vol0 + c);
    };
    if (fits && threadIdx.x == 0)
        for (int c = 0; c < min(C, NSLOT - 1); ++c) issue(c);

    const uint32_t GX = (uint32_t)Yi * Zi, GY = (uint32_t)Zi;
    // backward: the upstream gradient of channel c + 1 is fetched while channel c is processed
    float gcur[MT_X], gnext[MT_X];
#pragma unroll
    for (int i = 0; i < MT_X; ++i) {
        gcur[i] = gnext[i] = 0.f;
        if (BWD && i < nx && !dead[i]) gcur[i] = __ldg(gout + (size_t)vol0 * N + vox0 + i * XS);
    }
    for (int c = 0; c < C; ++c) {
        const float *ic = img + ((size_t)vol0 + c) * Ni;
        if (BWD && c + 1 < C) {
#pragma unroll
            for (int i = 0; i < MT_X; ++i)
                if (i < nx && !dead[i]) gnext[i] = __ldg(gout + ((size_t)vol0 + c + 1) * N + vox0 + i * XS);
        }
        const float *q = ring + (size_t)(c % NSLOT) * CS;
        if (fits) mbar_wait(&bar[c % NSLOT], (uint32_t)((c / NSLOT) & 1));
#pragma unroll
        for (int i = 0; i < MT_X; ++i) {
            if (i >= nx) break;
            float val[8];
            if (fits) {
                const float *p = q + base[i];
                val[0] = p[0]; val[1] = p[1]; val[2] = p[PY]; val[3] = p[PY + 1];
                val[4] = p[PX]; val[5] = p[PX + 1]; val[6] = p[PX + PY]; val[7] = p[PX + PY + 1];
            } else {
                gather8(ic + (uint32_t)base[i], GY, GX, 1u, val);
            }
            if (!BWD) {
                const float r = tri_accumulate(w[i], val);
                out[((size_t)vol0 + c) * N + vox0 + i * XS] = dead[i] ? fill : r;
            } else {
                const float g = gcur[i];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(g, val[k], acc[i][k]);
            }
        }
        if (BWD) {
#pragma unroll
            for (int i = 0; i < MT_X; ++i) gcur[i] = gnext[i];
        }
        if (fits) {
            __syncthreads();                          // slot (c % NSLOT) is free again
            if (threadIdx.x == 0 && c + NSLOT - 1 < C) issue(c + NSLOT - 1);
        }
    }

    if (BWD) {
        float *gf = gfield + (size_t)blockIdx.z * 3 * N;
#pragma unroll
        for (int i = 0; i < MT_X; ++i) {
            if (i >= nx) break;
            // d w_k / d loc: sign(corner) * inb * (other two weights), corner order (x,y,z) = k>>2, (k>>1)&1, k&1
            const float ibx = (l[0][i] >= 0.f && l[0][i] <= mxf) ? 1.f : 0.f;
            const float iby = (l[1][i] >= 0.f && l[1][i] <= myf) ? 1.f : 0.f;
            const float ibz = (l[2][i] >= 0.f && l[2][i] <= mzf) ? 1.f : 0.f;
            const float wx[2] = {ax[i].w0, ax[i].w1}, wy[2] = {ay[i].w0, ay[i].w1}, wz[2] = {az[i].w0, az[i].w1};
            float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int a = k >> 2, b = (k >> 1) & 1, d = k & 1;
                gx = fmaf((a ? ibx : -ibx) * wy[b] * wz[d], acc[i][k], gx);
                gy = fmaf(wx[a] * (b ? iby : -iby) * wz[d], acc[i][k], gy);
                gz = fmaf(wx[a] * wy[b] * (d ? ibz : -ibz), acc[i][k], gz);
            }
            const uint32_t vox = vox0 + i * XS;
            gf[vox] = gx; gf[N + vox] = gy; gf[2 * (size_t)N + vox] = gz;
        }
    }
}

template <bool BWD>
static int launch_mc(const float *img, const float *field, float *out, const float *gout, float *gfield, int B,
                     int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, float fill, cudaStream_t st) {
    constexpr int BX = 8, BY = 14, BZ = 48, NSLOT = 4;
    static const bool off = getenv("DFM_NO_BRICK") != nullptr;
    if (off || !tma_planar_ok(img, Xi, Yi, Zi) || Xi < 2 || Yi < 2 || Zi < 4) return DFM_EUNSUPPORTED;
    CUtensorMap tmap;
    if (!encode_planar_map(&tmap, img, B * C, Xi, Yi, Zi, BX, BY, BZ, 1)) return DFM_EUNSUPPORTED;
    constexpr size_t smem = (size_t)NSLOT * BX * BY * BZ * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_mc_brick<BX, BY, BZ, NSLOT, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_mc_brick smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int nzt = (Z + MT_Z - 1) / MT_Z, nyt = (Y + MT_Y - 1) / MT_Y, nxt = (X + MT_X - 1) / MT_X;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    k_warp_mc_brick<BX, BY, BZ, NSLOT, BWD><<<grid, block, smem, st>>>(tmap, img, field, out, gout, gfield, C, Xi, Yi, Zi,
                                                                         X, Y, Z, has_fill, fill, nzt);
    return check_launch(BWD ? "k_warp_mc_brick(bwd)" : "k_warp_mc_brick(fwd)");
}

int launch_warp_mc_fwd(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi, int X,
                       int Y, int Z, int has_fill, float fill, cudaStream_t st) {
    return launch_mc<false>(img, field, out, nullptr, nullptr, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, fill, st);
}
int launch_warp_mc_bwd_field(const float *gout, const float *img, const float *field, float *gfield, int B, int C,
                             int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, cudaStream_t st) {
    return launch_mc<true>(img, field, nullptr, gout, gfield, B, C, Xi, Yi, Zi, X, Y, Z, has_fill, 0.f, st);
}

}  // namespace dfm
