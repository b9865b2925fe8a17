// TMA-staged shared-memory bricks for the gather-heavy kernels (planar fp32):
//   k_ss_brick   : out = scale*own + interp(scale*src, p + scale*own)   (SS step, compose)
//   k_warp_brick : out = interp(img, p + field)                         (one-channel image warp)
//
// A CTA owns a tile of TX x TY x TZ = 8 x 8 x 32 output voxels (warp = one y row, lane = z,
// each thread walks the 8 x planes).  Pass 1 loads the tile's own vectors (coalesced), forms
// the sample locations and block-reduces the integer bounding box of their corner indices.
// One elected thread then issues a single 4-D TMA box load whose ORIGIN is that bounding
// box's corner -- the brick follows the displacement, so its size only has to cover the tile
// plus the local deformation, not the displacement magnitude.  Pass 2 gathers the corner
// values from shared memory: with the lower corner addressed as i1 - 1 (dfm_common.cuh) the 8
// corners are ONE base address plus compile-time offsets, i.e. 24 LDS with immediates and ~4
// address instructions per voxel (the direct kernel spends ~120 on 64-bit addressing).
// The brick's z pitch is 64 floats: every row and plane starts on bank 0, so lanes whose
// x/y corner rows differ still hit distinct banks.  Tiles whose box does not fit the brick
// (strong local deformation) take a per-thread checked path with global-memory gathers, so
// the result never depends on the brick size.  TMA zero-fills out-of-volume elements; they
// are only ever multiplied by an exactly-zero weight.
//
// Measured on sm_100a (scripts/tma_probe.cu): the innermost TMA coordinate must be a multiple
// of 16 bytes, otherwise the copy raises an illegal-instruction fault -> the box origin is
// aligned down to 4 floats in z.
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

namespace dfm {

constexpr int TY = 8, TZ = 32;

// block-wide bounding box of the corner indices: lower corner i1 - 1, upper corner i1.
// i1 is monotone in the clipped location, so each thread tracks min/max of its CLIPPED locations in
// float and converts only the six extremes; every lane then issues the shared-memory atomics (the
// compiler aggregates them per warp with CREDUX + one elected ATOMS -- doing the warp reduction by
// hand as well makes it reduce twice).
struct BoxReduce {
    float mn[3], mx[3];
    __device__ __forceinline__ void first(float cx, float cy, float cz) {
        mn[0] = mx[0] = cx; mn[1] = mx[1] = cy; mn[2] = mx[2] = cz;
    }
    __device__ __forceinline__ void add(float cx, float cy, float cz) {
        mn[0] = fminf(mn[0], cx); mn[1] = fminf(mn[1], cy); mn[2] = fminf(mn[2], cz);
        mx[0] = fmaxf(mx[0], cx); mx[1] = fmaxf(mx[1], cy); mx[2] = fmaxf(mx[2], cz);
    }
    // s_min/s_max: shared int[3] initialised to INT_MAX / INT_MIN before the preceding barrier;
    // they receive min(i1) and max(i1)
    __device__ __forceinline__ void commit(int *s_min, int *s_max, int mxi, int myi, int mzi) {
        const int mi[3] = {mxi, myi, mzi};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&s_min[d], axis_clipped_i1(mn[d], mi[d]));
            atomicMax(&s_max[d], axis_clipped_i1(mx[d], mi[d]));
        }
    }
};

// =========================================================================================
// field self / cross warp
// =========================================================================================
template <int TX, int BX, int BY, int BZ, bool SCALED>
__global__ void __launch_bounds__(256)
k_ss_brick(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src,
           const float *__restrict__ own, float *__restrict__ out, int Xs, int Ys, int Zs, int X, int Y,
           int Z, float scale, FastDiv nzt, const float *__restrict__ bound, float bscale) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *brick = reinterpret_cast<float *>(smem_raw);   // [3][BX][BY][BZ]
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_min[3], s_max[3];
    constexpr int PX = BY * BZ, PY = BZ, CS = BX * BY * BZ;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int yt = (int)fast_div(blockIdx.x, nzt), zt = (int)blockIdx.x - yt * (int)nzt.d;
    const int z = zt * TZ + lane, y = yt * TY + warp, x0 = blockIdx.y * TX;
    const bool ok_yz = (z < Z) && (y < Y);
    // threads past the volume edge shadow the nearest voxel inside (loads and the bounding box run
    // unpredicated; only their stores are masked)
    const int zc = min(z, Z - 1), yc = min(y, Y - 1);
    const int nx = min(TX, X - x0);                       // CTA-uniform, >= 1
    const uint32_t N = (uint32_t)X * Y * Z, Ns = (uint32_t)Xs * Ys * Zs;
    const float *ownb = own + (size_t)blockIdx.z * 3 * N;
    float *outb = out + (size_t)blockIdx.z * 3 * N;
    const int mxi = Xs - 1, myi = Ys - 1, mzi = Zs - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fy = (float)yc, fz = (float)zc, fx0 = (float)x0;
    const uint32_t vox0 = ((uint32_t)x0 * Y + yc) * Z + zc, XS = (uint32_t)Y * Z;
    const bool same_grid = (Xs == X) && (Ys == Y) && (Zs == Z);

    // Static halo: when the caller can bound the displacements of this batch item (|v| < HS voxels --
    // the early scaling-and-squaring steps; `bound[b] * bscale` is that bound), the corners of the
    // tile lie inside [tile - HS, tile + HS], so the brick is requested before the own vectors are
    // even loaded and the bounding-box reduction (and its barrier) disappears.
    constexpr int HS_XY = (BX - TX) / 2 < (BY - TY) / 2 ? (BX - TX) / 2 : (BY - TY) / 2;
    constexpr int HS_Z = BZ - TZ - 5 < 4 ? BZ - TZ - 5 : 4;      // origin z0 - 4 (16-byte aligned), upper corner z0 + 31 + HS + 1
    constexpr int HS = HS_XY < HS_Z ? HS_XY : HS_Z;
    const bool stat = HS >= 1 && bound != nullptr && same_grid &&
                      __ldg(bound + blockIdx.z) * bscale < (float)HS * 0.999f;       // CTA-uniform

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        if (stat) {
            mbar_expect_tx(&bar, 3u * CS * sizeof(float));
            tma_load_4d(brick, &tmap, &bar, zt * TZ - 4, yt * TY - HS, x0 - HS, (int)blockIdx.z * 3);
        } else {
            s_min[0] = s_min[1] = s_min[2] = INT_MAX;
            s_max[0] = s_max[1] = s_max[2] = INT_MIN;
        }
    }
    __syncthreads();

    // ---- pass 1: own vectors, sample locations, bounding box --------------------------------
    float v[3][TX];
    int ox, oy, oz;
    bool fits;
    if (stat) {
        // the tile lies inside its own static brick: the own vectors come from shared memory too
        ox = x0 - HS; oy = yt * TY - HS; oz = zt * TZ - 4;             // brick origin = lower corner index
        fits = true;
        mbar_wait(&bar, 0);
        const float *qo = brick + (HS * PX + (yc - oy) * PY + (zc - oz));
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            const int ic = min(i, nx - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c][i] = qo[c * CS + ic * PX];
                if (SCALED) v[c][i] = __fmul_rn(scale, v[c][i]);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            const int ic = min(i, nx - 1);
            const float *pv = ownb + vox0 + ic * XS;
            v[0][i] = __ldg(pv);
            v[1][i] = __ldg(pv + N);
            v[2][i] = __ldg(pv + 2 * (size_t)N);
        }
        BoxReduce box;
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            if (SCALED) {
                v[0][i] = __fmul_rn(scale, v[0][i]); v[1][i] = __fmul_rn(scale, v[1][i]); v[2][i] = __fmul_rn(scale, v[2][i]);
            }
            const float fx = fx0 + (float)min(i, nx - 1);
            const float cx = axis_clip(__fadd_rn(fx, v[0][i]), mxf), cy = axis_clip(__fadd_rn(fy, v[1][i]), myf),
                        cz = axis_clip(__fadd_rn(fz, v[2][i]), mzf);
            if (i == 0) box.first(cx, cy, cz); else box.add(cx, cy, cz);
        }
        box.commit(s_min, s_max, mxi, myi, mzi);
        __syncthreads();
        ox = s_min[0] - 1; oy = s_min[1] - 1; oz = (s_min[2] - 1) & ~3;           // lower corner index
        fits = (s_max[0] - ox < BX) && (s_max[1] - oy < BY) && (s_max[2] - oz < BZ);
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, 3u * CS * sizeof(float));
            tma_load_4d(brick, &tmap, &bar, oz, oy, ox, (int)blockIdx.z * 3);
        }
    }
    mbar_wait(&bar, 0);

    // ---- pass 2 ------------------------------------------------------------------------------
    // lower-corner offset in the brick = i1x*PX + i1y*PY + i1z + cbase
    const int cbase = -((ox + 1) * PX + (oy + 1) * PY + (oz + 1));
    if (fits) {
        float *o0 = outb + vox0, *o1 = o0 + N, *o2 = o1 + N;
        auto voxel = [&](int i) {
            const float v0 = v[0][i], v1 = v[1][i], v2 = v[2][i];
            const AxisF ax = axis_fast(__fadd_rn(fx0 + (float)i, v0), mxf, mxi);
            const AxisF ay = axis_fast(__fadd_rn(fy, v1), myf, myi);
            const AxisF az = axis_fast(__fadd_rn(fz, v2), mzf, mzi);
            float w[8];
            tri_weights(ax, ay, az, w);
            const float *q = brick + (ax.i1 * PX + ay.i1 * PY + az.i1 + cbase);
            float a[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float val[8] = {q[c * CS], q[c * CS + 1], q[c * CS + PY], q[c * CS + PY + 1],
                                      q[c * CS + PX], q[c * CS + PX + 1], q[c * CS + PX + PY], q[c * CS + PX + PY + 1]};
                a[c] = tri_accumulate(w, val);
            }
            if (SCALED) { a[0] = __fmul_rn(scale, a[0]); a[1] = __fmul_rn(scale, a[1]); a[2] = __fmul_rn(scale, a[2]); }
            o0[i * XS] = __fadd_rn(v0, a[0]);
            o1[i * XS] = __fadd_rn(v1, a[1]);
            o2[i * XS] = __fadd_rn(v2, a[2]);
        };
        if (!ok_yz) return;
        if (nx == TX) {                               // interior tile: no predication in the loop
#pragma unroll
            for (int i = 0; i < TX; ++i) voxel(i);
        } else {
#pragma unroll
            for (int i = 0; i < TX; ++i)
                if (i < nx) voxel(i);
        }
    } else {
        if (!ok_yz) return;
        const float *srcb = src + (size_t)blockIdx.z * 3 * Ns;
        const uint32_t GX = (uint32_t)Ys * Zs, GY = (uint32_t)Zs;
        for (int i = 0; i < nx; ++i) {
            float v0, v1, v2;                         // dynamic index -> select (keeps v[] in registers)
#pragma unroll
            for (int k = 0; k < TX; ++k)
                if (k == i) { v0 = v[0][k]; v1 = v[1][k]; v2 = v[2][k]; }
            const AxisF ax = axis_fast(__fadd_rn(fx0 + (float)i, v0), mxf, mxi);
            const AxisF ay = axis_fast(__fadd_rn(fy, v1), myf, myi);
            const AxisF az = axis_fast(__fadd_rn(fz, v2), mzf, mzi);
            float w[8];
            tri_weights(ax, ay, az, w);
            float a[3];
            if ((ax.i1 - ox < BX) && (ay.i1 - oy < BY) && (az.i1 - oz < BZ)) {
                const float *q = brick + (ax.i1 * PX + ay.i1 * PY + az.i1 + cbase);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float val[8] = {q[c * CS], q[c * CS + 1], q[c * CS + PY], q[c * CS + PY + 1],
                                          q[c * CS + PX], q[c * CS + PX + 1], q[c * CS + PX + PY], q[c * CS + PX + PY + 1]};
                    a[c] = tri_accumulate(w, val);
                }
            } else {
                const float *g = srcb + ((uint32_t)(ax.i1 - 1) * GX + (uint32_t)(ay.i1 - 1) * GY + (uint32_t)(az.i1 - 1));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float *gc = g + (size_t)c * Ns;
                    const float val[8] = {__ldg(gc), __ldg(gc + 1), __ldg(gc + GY), __ldg(gc + GY + 1),
                                          __ldg(gc + GX), __ldg(gc + GX + 1), __ldg(gc + GX + GY), __ldg(gc + GX + GY + 1)};
                    a[c] = tri_accumulate(w, val);
                }
            }
            if (SCALED) { a[0] = __fmul_rn(scale, a[0]); a[1] = __fmul_rn(scale, a[1]); a[2] = __fmul_rn(scale, a[2]); }
            const uint32_t vox = vox0 + i * XS;
            outb[vox] = __fadd_rn(v0, a[0]);
            outb[N + vox] = __fadd_rn(v1, a[1]);
            outb[2 * (size_t)N + vox] = __fadd_rn(v2, a[2]);
        }
    }
}

// -----------------------------------------------------------------------------------------
// First scaling-and-squaring step of an inference call: the svf arrives channels-last
// [B][X][Y][Z][3] and is scaled by 2^-nsteps, so its displacements are almost always far below
// one voxel.  The kernel is OPTIMISTIC: it requests the tile's static-halo brick straight from
// the channels-last tensor (one TMA box {3*BZ, BY, BX}; a corner is three adjacent floats, lanes
// stride by 3 floats = conflict-free), reads its own vectors from that brick, and only if some
// voxel of the CTA turns out to move by a voxel or more does the CTA fall back to per-voxel
// global gathers.  Output planar; max |out| per batch item goes to `absmax` (static-halo bound of
// the following steps).  Same arithmetic as k_field_warp_add / k_ss_brick<SCALED>.
// -----------------------------------------------------------------------------------------
template <int TX, int BX, int BY, int BZ>
__global__ void __launch_bounds__(256)
k_ss_first_cl(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ svf, float *__restrict__ out, int X,
              int Y, int Z, float scale, FastDiv nzt, float *__restrict__ absmax) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float *brick = reinterpret_cast<const float *>(smem_raw);   // [BX][BY][BZ][3]
    __shared__ __align__(8) uint64_t bar;
    constexpr int PX = BY * BZ, PY = BZ, CS = BX * BY * BZ;
    constexpr int HS = 1;
    static_assert(BX >= TX + 2 * HS && BY >= TY + 2 * HS && BZ >= TZ + 4 + 1 + HS, "static halo does not fit the box");

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int yt = (int)fast_div(blockIdx.x, nzt), zt = (int)blockIdx.x - yt * (int)nzt.d;
    const int z = zt * TZ + lane, y = yt * TY + warp, x0 = blockIdx.y * TX;
    const bool ok_yz = (z < Z) && (y < Y);
    const int zc = min(z, Z - 1), yc = min(y, Y - 1);
    const int nx = min(TX, X - x0);
    const uint32_t N = (uint32_t)X * Y * Z;
    const float *srcb = svf + (size_t)blockIdx.z * 3 * N;
    float *outb = out + (size_t)blockIdx.z * 3 * N;
    const int mxi = X - 1, myi = Y - 1, mzi = Z - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fy = (float)yc, fz = (float)zc, fx0 = (float)x0;
    const uint32_t vox0 = ((uint32_t)x0 * Y + yc) * Z + zc, XS = (uint32_t)Y * Z;
    const int ox = x0 - HS, oy = yt * TY - HS, oz = zt * TZ - 4;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 3u * CS * sizeof(float));
        tma_load_4d(smem_raw, &tmap, &bar, 3 * oz, oy, ox, (int)blockIdx.z);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    float v[3][TX];
    bool big = false;
    {
        const float *qo = brick + 3 * (HS * PX + (yc - oy) * PY + (zc - oz));
#pragma unroll
        for (int i = 0; i < TX; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[c][i] = __fmul_rn(scale, qo[3 * min(i, nx - 1) * PX + c]);
                big |= !(fabsf(v[c][i]) < 0.999f * (float)HS);          // also true for NaN
            }
    }
    const bool fallback = __syncthreads_or(big) != 0;                   // CTA-uniform

    const int cbase = -((ox + 1) * PX + (oy + 1) * PY + (oz + 1));
    const uint32_t GX = (uint32_t)Y * Z, GY = (uint32_t)Z;
    float am = 0.f;
    if (ok_yz) {
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            if (i >= nx) break;
            const float v0 = v[0][i], v1 = v[1][i], v2 = v[2][i];
            const AxisF ax = axis_fast(__fadd_rn(fx0 + (float)i, v0), mxf, mxi);
            const AxisF ay = axis_fast(__fadd_rn(fy, v1), myf, myi);
            const AxisF az = axis_fast(__fadd_rn(fz, v2), mzf, mzi);
            float w[8], a[3];
            tri_weights(ax, ay, az, w);
            if (!fallback) {
                const float *q = brick + 3 * (ax.i1 * PX + ay.i1 * PY + az.i1 + cbase);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float val[8] = {q[c], q[c + 3], q[c + 3 * PY], q[c + 3 * (PY + 1)],
                                          q[c + 3 * PX], q[c + 3 * (PX + 1)], q[c + 3 * (PX + PY)], q[c + 3 * (PX + PY + 1)]};
                    a[c] = tri_accumulate(w, val);
                }
            } else {
                const float *g = srcb + 3 * (size_t)((uint32_t)(ax.i1 - 1) * GX + (uint32_t)(ay.i1 - 1) * GY + (uint32_t)(az.i1 - 1));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float val[8];
                    gather8(g + c, 3u * GY, 3u * GX, 3u, val);
                    a[c] = tri_accumulate(w, val);
                }
            }
            const float r0 = __fadd_rn(v0, __fmul_rn(scale, a[0]));
            const float r1 = __fadd_rn(v1, __fmul_rn(scale, a[1]));
            const float r2 = __fadd_rn(v2, __fmul_rn(scale, a[2]));
            am = absmax_fold(absmax_fold(absmax_fold(am, r0), r1), r2);
            const uint32_t vox = vox0 + i * XS;
            outb[vox] = r0; outb[N + vox] = r1; outb[2 * (size_t)N + vox] = r2;
        }
    }
    if (absmax) block_absmax_commit(am, absmax + blockIdx.z);          // uniform branch, every thread arrives
}

// =========================================================================================
// one-channel image warp (linear)
// =========================================================================================
// FMODE 3 = the same fusion evaluated SEPARABLY (default build): the tile's coarse box arrives by one TMA
// load issued at kernel start; a thread reduces each coarse plane it crosses once to its (y,z)-bilinear
// value (4 taps x 3 components, weights fixed per thread and carrying the vector scale) and every voxel of
// its x walk is one lerp between two held planes -- ~25 instructions per voxel instead of ~90, 11.5 KB of
// shared memory instead of a float4 box.  Differs from the reference summation order by a few ulp.
// FMODE: 0 = planar field, 1 = channels-last field, 2 = FUSED RescaleTransform: `field` is the
// coarse planar field [B][3][Xh][Yh][Zh]; the displacement of every output voxel is resampled on
// the fly from a shared-memory copy of the coarse box the tile touches (same arithmetic as
// k_resize3_smem: pre-scaled values, reference corner order), so the full-resolution field is
// never written to or read from HBM.
struct UpsampleArgs {
    const float *cx, *cy, *cz;      // sample coordinates of the output grid on the coarse grid
    int Xh, Yh, Zh;
    float pre;                      // vector scale applied before resampling (zoom factor)
    int cap;                        // capacity (float4) of the coarse box in shared memory
};

// coarse box of FMODE 3 (TMA, one box per CTA): covers the 4 x 8 x 32 tile for zoom factors >= 2
constexpr int HBX = 5, HBY = 8, HBZ = 24, HCS = HBX * HBY * HBZ;

template <int TX, int BX, int BY, int BZ, int FMODE, bool HF>
__global__ void __launch_bounds__(256)
k_warp_brick(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_h, const float *__restrict__ img,
             const float *__restrict__ field, float *__restrict__ out, int Xi, int Yi, int Zi, int X, int Y,
             int Z, float fill, FastDiv nzt, UpsampleArgs up) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *brick = reinterpret_cast<float *>(smem_raw);   // [BX][BY][BZ]
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_min[3], s_max[3];
    constexpr int PX = BY * BZ, PY = BZ, CS = BX * BY * BZ;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int yt = (int)fast_div(blockIdx.x, nzt), zt = (int)blockIdx.x - yt * (int)nzt.d;
    const int z = zt * TZ + lane, y = yt * TY + warp, x0 = blockIdx.y * TX;
    const bool ok_yz = (z < Z) && (y < Y);
    // threads past the volume edge shadow the nearest voxel inside (loads and the bounding box run
    // unpredicated; only their stores are masked)
    const int zc = min(z, Z - 1), yc = min(y, Y - 1);
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    float *outb = out + (size_t)blockIdx.z * N;
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fy = (float)yc, fz = (float)zc, fx0 = (float)x0;
    const uint32_t vox0 = ((uint32_t)x0 * Y + yc) * Z + zc, XS = (uint32_t)Y * Z;
    const int nx = min(TX, X - x0);                       // CTA-uniform, >= 1

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        s_min[0] = s_min[1] = s_min[2] = INT_MAX;
        s_max[0] = s_max[1] = s_max[2] = INT_MIN;
    }
    __syncthreads();

    float l[3][TX];                                   // displacements, then CLIPPED sample locations
    if (FMODE == 3) {
        __shared__ __align__(8) uint64_t bar_h;
        float *hbox = brick + CS;                     // [3][HBX][HBY][HBZ]
        const int Xh = up.Xh, Yh = up.Yh, Zh = up.Zh;
        const float hxf = (float)(Xh - 1), hyf = (float)(Yh - 1), hzf = (float)(Zh - 1);
        // coarse box origin of this tile (tables are non-decreasing); z origin 16-byte aligned for TMA
        const int bx0 = axis_fast_i1(__ldg(up.cx + x0), hxf, Xh - 1) - 1;
        const int by0 = axis_fast_i1(__ldg(up.cy + yt * TY), hyf, Yh - 1) - 1;
        const int bz0 = (axis_fast_i1(__ldg(up.cz + zt * TZ), hzf, Zh - 1) - 1) & ~3;
        if (threadIdx.x == 0) {
            mbar_init(&bar_h, 1);
            mbar_expect_tx(&bar_h, (uint32_t)(3 * HCS * sizeof(float)));
            tma_load_4d(hbox, &tmap_h, &bar_h, bz0, by0, bx0, (int)blockIdx.z * 3);
        }
        const AxisF ay = axis_fast(__ldg(up.cy + yc), hyf, Yh - 1);
        const AxisF az = axis_fast(__ldg(up.cz + zc), hzf, Zh - 1);
        const float w4[4] = {up.pre * ay.w0 * az.w0, up.pre * ay.w0 * az.w1, up.pre * ay.w1 * az.w0, up.pre * ay.w1 * az.w1};
        const float *h0 = hbox + (ay.i1 - 1 - by0) * HBZ + (az.i1 - 1 - bz0);
        AxisF axs[TX];
#pragma unroll
        for (int i = 0; i < TX; ++i) axs[i] = axis_fast(__ldg(up.cx + x0 + min(i, nx - 1)), hxf, Xh - 1);
        __syncthreads();                              // bar_h initialised
        mbar_wait(&bar_h, 0);
        float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
        int have = -1;                                // relative index of the coarse plane held in `hi`
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            const int need = axs[i].i1 - bx0;         // relative index of the upper coarse plane (CTA-uniform)
            while (have < need) {
                ++have;
                const float *q = h0 + have * (HBY * HBZ);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    lo[c] = hi[c];
                    const float *qc = q + c * HCS;
                    hi[c] = fmaf(w4[3], qc[HBZ + 1], fmaf(w4[2], qc[HBZ], fmaf(w4[1], qc[1], w4[0] * qc[0])));
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) l[c][i] = fmaf(axs[i].w1, hi[c], axs[i].w0 * lo[c]);
        }
    } else if (FMODE == 2) {
        float4 *hbox = reinterpret_cast<float4 *>(smem_raw + (size_t)CS * sizeof(float));
        const int Xh = up.Xh, Yh = up.Yh, Zh = up.Zh;
        const uint32_t Nh = (uint32_t)Xh * Yh * Zh;
        const float *hb = field + (size_t)blockIdx.z * 3 * Nh;
        const int jx1 = min(x0 + TX, X) - 1, jy0 = yt * TY, jy1 = min(jy0 + TY, Y) - 1, jz0 = zt * TZ, jz1 = min(jz0 + TZ, Z) - 1;
        // coarse box of the tile: [i1(first) - 1, i1(last)] per axis (tables are non-decreasing)
        const int bx0 = axis_fast_i1(__ldg(up.cx + x0), (float)(Xh - 1), Xh - 1) - 1, bx1 = axis_fast_i1(__ldg(up.cx + jx1), (float)(Xh - 1), Xh - 1);
        const int by0 = axis_fast_i1(__ldg(up.cy + jy0), (float)(Yh - 1), Yh - 1) - 1, by1 = axis_fast_i1(__ldg(up.cy + jy1), (float)(Yh - 1), Yh - 1);
        const int bz0 = axis_fast_i1(__ldg(up.cz + jz0), (float)(Zh - 1), Zh - 1) - 1, bz1 = axis_fast_i1(__ldg(up.cz + jz1), (float)(Zh - 1), Zh - 1);
        const int nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1, nbz = bz1 - bz0 + 1;
        const bool staged = nbx * nby * nbz <= up.cap;                   // CTA-uniform
        if (staged) {
            const int total = nbx * nby * nbz;
            for (int t = threadIdx.x; t < total; t += 256) {
                const int bz = t % nbz, q = t / nbz, by = q % nby, bx = q / nby;
                const uint32_t o = ((uint32_t)(bx0 + bx) * Yh + (by0 + by)) * Zh + (bz0 + bz);
                hbox[t] = make_float4(__fmul_rn(up.pre, __ldg(hb + o)), __fmul_rn(up.pre, __ldg(hb + Nh + o)),
                                      __fmul_rn(up.pre, __ldg(hb + 2 * (size_t)Nh + o)), 0.f);
            }
        }
        __syncthreads();
        const AxisF ay = axis_fast(__ldg(up.cy + yc), (float)(Yh - 1), Yh - 1);
        const AxisF az = axis_fast(__ldg(up.cz + zc), (float)(Zh - 1), Zh - 1);
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            l[0][i] = l[1][i] = l[2][i] = 0.f;
            if (i < nx) {
                const AxisF ax = axis_fast(__ldg(up.cx + x0 + i), (float)(Xh - 1), Xh - 1);
                float w[8], v0[8], v1[8], v2[8];
                tri_weights(ax, ay, az, w);
                if (staged) {
                    const int sy = nbz, sx = nby * nbz;
                    const float4 *q = hbox + (((ax.i1 - 1 - bx0) * nby + (ay.i1 - 1 - by0)) * nbz + (az.i1 - 1 - bz0));
                    const int o[8] = {0, 1, sy, sy + 1, sx, sx + 1, sx + sy, sx + sy + 1};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 t = q[o[k]];
                        v0[k] = t.x; v1[k] = t.y; v2[k] = t.z;
                    }
                } else {
                    const uint32_t gy = (uint32_t)Zh, gx = (uint32_t)Yh * Zh;
                    const float *g = hb + ((uint32_t)(ax.i1 - 1) * gx + (uint32_t)(ay.i1 - 1) * gy + (uint32_t)(az.i1 - 1));
                    gather8(g, gy, gx, 1u, v0);
                    gather8(g + Nh, gy, gx, 1u, v1);
                    gather8(g + 2 * (size_t)Nh, gy, gx, 1u, v2);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { v0[k] = __fmul_rn(up.pre, v0[k]); v1[k] = __fmul_rn(up.pre, v1[k]); v2[k] = __fmul_rn(up.pre, v2[k]); }
                }
                l[0][i] = tri_accumulate(w, v0);
                l[1][i] = tri_accumulate(w, v1);
                l[2][i] = tri_accumulate(w, v2);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            const uint32_t vox = vox0 + min(i, nx - 1) * XS;
            if (FMODE == 1) {
                l[0][i] = __ldg(fb + (size_t)vox * 3); l[1][i] = __ldg(fb + (size_t)vox * 3 + 1); l[2][i] = __ldg(fb + (size_t)vox * 3 + 2);
            } else {
                l[0][i] = __ldg(fb + vox); l[1][i] = __ldg(fb + N + vox); l[2][i] = __ldg(fb + 2 * (size_t)N + vox);
            }
        }
    }
    BoxReduce box;
    uint32_t oob = 0;                                 // HF: voxels sampling outside the image
#pragma unroll
    for (int i = 0; i < TX; ++i) {
        const float lx = __fadd_rn(fx0 + (float)min(i, nx - 1), l[0][i]);
        const float ly = __fadd_rn(fy, l[1][i]);
        const float lz = __fadd_rn(fz, l[2][i]);
        if (HF && (lx < 0.f || lx > mxf || ly < 0.f || ly > myf || lz < 0.f || lz > mzf)) oob |= 1u << i;
        l[0][i] = axis_clip(lx, mxf); l[1][i] = axis_clip(ly, myf); l[2][i] = axis_clip(lz, mzf);
        if (i == 0) box.first(l[0][i], l[1][i], l[2][i]); else box.add(l[0][i], l[1][i], l[2][i]);
    }
    box.commit(s_min, s_max, mxi, myi, mzi);
    __syncthreads();
    const int ox = s_min[0] - 1, oy = s_min[1] - 1, oz = (s_min[2] - 1) & ~3;     // lower corner index
    const bool fits = (s_max[0] - ox < BX) && (s_max[1] - oy < BY) && (s_max[2] - oz < BZ);

    if (fits) {                                       // a box that does not fit is not worth staging
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)(CS * sizeof(float)));
            tma_load_4d(brick, &tmap, &bar, oz, oy, ox, (int)blockIdx.z);
        }
        mbar_wait(&bar, 0);
    }
    if (!ok_yz) return;

    const int cbase = -((ox + 1) * PX + (oy + 1) * PY + (oz + 1));
    if (fits) {
#pragma unroll
        for (int i = 0; i < TX; i += 2) {
            if (i >= nx) break;
            const bool hasB = i + 1 < nx;             // (the shadow of voxel i otherwise)
            const AxisF ax = axis_from_clipped(l[0][i], mxi), ay = axis_from_clipped(l[1][i], myi), az = axis_from_clipped(l[2][i], mzi);
            const AxisF bx = axis_from_clipped(l[0][i + 1], mxi), by = axis_from_clipped(l[1][i + 1], myi), bz = axis_from_clipped(l[2][i + 1], mzi);
            float wA[8], wB[8];
            tri_weights_pair(ax, ay, az, bx, by, bz, wA, wB);
            const float *qa = brick + (ax.i1 * PX + ay.i1 * PY + az.i1 + cbase);
            const float *qb = brick + (bx.i1 * PX + by.i1 * PY + bz.i1 + cbase);
            const float va[8] = {qa[0], qa[1], qa[PY], qa[PY + 1], qa[PX], qa[PX + 1], qa[PX + PY], qa[PX + PY + 1]};
            const float vb[8] = {qb[0], qb[1], qb[PY], qb[PY + 1], qb[PX], qb[PX + 1], qb[PX + PY], qb[PX + PY + 1]};
            float ra, rb;
            tri_accumulate_pair(wA, wB, va, vb, ra, rb);
            if (HF) {
                if (oob & (1u << i)) ra = fill;
                if (oob & (2u << i)) rb = fill;
            }
            outb[vox0 + i * XS] = ra;
            if (hasB) outb[vox0 + (i + 1) * XS] = rb;
        }
    } else {
        const float *ib = img + (size_t)blockIdx.z * Ni;
        const uint32_t GX = (uint32_t)Yi * Zi, GY = (uint32_t)Zi;
#pragma unroll
        for (int i = 0; i < TX; ++i) {
            if (i >= nx) break;
            const AxisF ax = axis_from_clipped(l[0][i], mxi), ay = axis_from_clipped(l[1][i], myi), az = axis_from_clipped(l[2][i], mzi);
            float w[8], val[8];
            tri_weights(ax, ay, az, w);
            gather8(ib + ((uint32_t)(ax.i1 - 1) * GX + (uint32_t)(ay.i1 - 1) * GY + (uint32_t)(az.i1 - 1)), GY, GX, 1u, val);
            float r = tri_accumulate(w, val);
            if (HF && (oob & (1u << i))) r = fill;
            outb[vox0 + i * XS] = r;
        }
    }
}

// -----------------------------------------------------------------------------------------
// One-channel linear image warp with a VARIABLE brick: 8 x 8 x 16 voxels per CTA (warp = two y rows x 16 z,
// each thread walks 4 x planes).
// On strongly deforming fields (bench: std-3 SVF, |grad u| ~ 0.23 at full resolution, |u| up to 17 voxels)
// the bounding box of a tile is dominated by its slant along the long z edge: a 4 x 8 x 32 tile needs
// 11 x 14 x 34 voxels in the median and fits the fixed 10 x 16 x 48 box of k_warp_brick in 28 % of the cases;
// any fixed box that fits >= 94 % of the tiles is >= 11 x the tile (L2 -> SM traffic).  Here the box is cut to
// the tile: the brick is loaded as ONE TMA BOX PER X PLANE, as many planes as the bounding box has, each
// {BZ, BY} with BY in {12, 16, 20, 24} and BZ in {24, 32} picked per CTA from eight tensor maps -- 97 % of the
// bench tiles fit 36 KB at 4.6 x the tile on average (the single 18 x 18 x 24 box: 76 % at 5.8 x).
// Two rows per warp cost a 2-way bank conflict on most gathers; the one-channel warp has 8 gathers per voxel
// and is nowhere near the shared-memory limit (the SS step, with 24, is).
// Same arithmetic as k_warp_brick (FMODE 0 / 1); tiles that do not fit gather from global memory.
// -----------------------------------------------------------------------------------------
struct WarpMaps {
    CUtensorMap m[4][2];            // [BY = 12 + 4 i][BZ = 24 + 8 j], box {BZ, BY, 1, 1}
};

template <bool FIELD_CL, bool HF>
__global__ void __launch_bounds__(256)
k_warp_brick_var(const __grid_constant__ WarpMaps maps, const float *__restrict__ img, const float *__restrict__ field,
                 float *__restrict__ out, int Xi, int Yi, int Zi, int X, int Y, int Z, float fill, FastDiv nzt, int cap_floats) {
    constexpr int TXC = 8, TYC = 8, TZC = 16, NX = 4;          // CTA tile; x planes per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *brick = reinterpret_cast<float *>(smem_raw);        // [ex][BY][BZ]
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_min[3], s_max[3];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int yt = (int)fast_div(blockIdx.x, nzt), zt = (int)blockIdx.x - yt * (int)nzt.d;
    const int z = zt * TZC + (lane & 15), y = yt * TYC + 2 * (warp & 3) + (lane >> 4), x0 = blockIdx.y * TXC + (warp >> 2) * NX;
    const bool ok_yz = (z < Z) && (y < Y);
    const int zc = min(z, Z - 1), yc = min(y, Y - 1);
    const uint32_t N = (uint32_t)X * Y * Z, Ni = (uint32_t)Xi * Yi * Zi;
    const float *fb = field + (size_t)blockIdx.z * 3 * N;
    float *outb = out + (size_t)blockIdx.z * N;
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fy = (float)yc, fz = (float)zc;
    const int nx = min(NX, X - x0);                            // per half-CTA; may be <= 0 past the volume edge
    const int x0c = min(x0, X - 1);
    const uint32_t XS = (uint32_t)Y * Z, vox0 = ((uint32_t)x0c * Y + yc) * Z + zc;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        s_min[0] = s_min[1] = s_min[2] = INT_MAX;
        s_max[0] = s_max[1] = s_max[2] = INT_MIN;
    }
    __syncthreads();

    float l[3][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const uint32_t vox = vox0 + (uint32_t)min(i, max(nx, 1) - 1) * XS;     // shadow the last valid plane
        if (FIELD_CL) {
            l[0][i] = __ldg(fb + (size_t)vox * 3); l[1][i] = __ldg(fb + (size_t)vox * 3 + 1); l[2][i] = __ldg(fb + (size_t)vox * 3 + 2);
        } else {
            l[0][i] = __ldg(fb + vox); l[1][i] = __ldg(fb + N + vox); l[2][i] = __ldg(fb + 2 * (size_t)N + vox);
        }
    }
    BoxReduce box;
    uint32_t oob = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const float lx = __fadd_rn((float)(x0c + min(i, max(nx, 1) - 1)), l[0][i]);
        const float ly = __fadd_rn(fy, l[1][i]);
        const float lz = __fadd_rn(fz, l[2][i]);
        if (HF && (lx < 0.f || lx > mxf || ly < 0.f || ly > myf || lz < 0.f || lz > mzf)) oob |= 1u << i;
        l[0][i] = axis_clip(lx, mxf); l[1][i] = axis_clip(ly, myf); l[2][i] = axis_clip(lz, mzf);
        if (i == 0) box.first(l[0][i], l[1][i], l[2][i]); else box.add(l[0][i], l[1][i], l[2][i]);
    }
    box.commit(s_min, s_max, mxi, myi, mzi);
    __syncthreads();
    const int ox = s_min[0] - 1, oy = s_min[1] - 1, oz = (s_min[2] - 1) & ~3;                  // lower corner index
    const int ex = s_max[0] - ox + 1, ey = s_max[1] - oy + 1, ez = s_max[2] - oz + 1;          // extents of the corner indices
    const int iy = max((ey + 3) / 4 - 3, 0), iz = ez <= 24 ? 0 : 1;                            // BY = 12 + 4 iy, BZ = 24 + 8 iz
    const int PY = 24 + 8 * iz, PX = (12 + 4 * iy) * PY;
    const bool fits = (iy <= 3) && (ez <= 32) && (ex * PX <= cap_floats);                       // CTA-uniform
    if (fits) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)(ex * PX * sizeof(float)));
            // one box per x plane from the map whose {BZ, BY} covers the bounding box (constant indices: the
            // tensor maps must be addressed in parameter space)
#define DFM_PLANES(I, J)                                                                                   \
    case (I) * 2 + (J):                                                                                    \
        for (int p = 0; p < ex; ++p) tma_load_4d(brick + p * PX, &maps.m[I][J], &bar, oz, oy, ox + p, (int)blockIdx.z); \
        break;
            switch (iy * 2 + iz) {
                DFM_PLANES(0, 0) DFM_PLANES(0, 1) DFM_PLANES(1, 0) DFM_PLANES(1, 1)
                DFM_PLANES(2, 0) DFM_PLANES(2, 1) DFM_PLANES(3, 0) DFM_PLANES(3, 1)
            }
#undef DFM_PLANES
        }
        mbar_wait(&bar, 0);
    }
    if (!ok_yz || nx <= 0) return;
    const int cbase = -((ox + 1) * PX + (oy + 1) * PY + (oz + 1));
    if (fits) {
#pragma unroll
        for (int i = 0; i < NX; i += 2) {
            if (i >= nx) break;
            const bool hasB = i + 1 < nx;
            const AxisF ax = axis_from_clipped(l[0][i], mxi), ay = axis_from_clipped(l[1][i], myi), az = axis_from_clipped(l[2][i], mzi);
            const AxisF bx = axis_from_clipped(l[0][i + 1], mxi), by = axis_from_clipped(l[1][i + 1], myi), bz = axis_from_clipped(l[2][i + 1], mzi);
            float wA[8], wB[8];
            tri_weights_pair(ax, ay, az, bx, by, bz, wA, wB);
            const float *qa = brick + (ax.i1 * PX + ay.i1 * PY + az.i1 + cbase), *qa1 = qa + PY, *qa2 = qa + PX, *qa3 = qa2 + PY;
            const float *qb = brick + (bx.i1 * PX + by.i1 * PY + bz.i1 + cbase), *qb1 = qb + PY, *qb2 = qb + PX, *qb3 = qb2 + PY;
            const float va[8] = {qa[0], qa[1], qa1[0], qa1[1], qa2[0], qa2[1], qa3[0], qa3[1]};
            const float vb[8] = {qb[0], qb[1], qb1[0], qb1[1], qb2[0], qb2[1], qb3[0], qb3[1]};
            float ra, rb;
            tri_accumulate_pair(wA, wB, va, vb, ra, rb);
            if (HF) {
                if (oob & (1u << i)) ra = fill;
                if (oob & (2u << i)) rb = fill;
            }
            outb[vox0 + i * XS] = ra;
            if (hasB) outb[vox0 + (i + 1) * XS] = rb;
        }
    } else {
        const float *ib = img + (size_t)blockIdx.z * Ni;
        const uint32_t GX = (uint32_t)Yi * Zi, GY = (uint32_t)Zi;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            if (i >= nx) break;
            const AxisF ax = axis_from_clipped(l[0][i], mxi), ay = axis_from_clipped(l[1][i], myi), az = axis_from_clipped(l[2][i], mzi);
            float w[8], val[8];
            tri_weights(ax, ay, az, w);
            gather8(ib + ((uint32_t)(ax.i1 - 1) * GX + (uint32_t)(ay.i1 - 1) * GY + (uint32_t)(az.i1 - 1)), GY, GX, 1u, val);
            float r = tri_accumulate(w, val);
            if (HF && (oob & (1u << i))) r = fill;
            outb[vox0 + i * XS] = r;
        }
    }
}

// ------------------------------- host side -----------------------------------------------
static bool brick_disabled() {
    static const bool d = getenv("DFM_NO_BRICK") != nullptr;   // debugging aid: force direct gathers
    return d;
}
static bool encode_map(CUtensorMap *tmap, const float *base, int nvol, int X, int Y, int Z, int bx, int by,
                       int bz, int bc) {
    return encode_planar_map(tmap, base, nvol, X, Y, Z, bx, by, bz, bc);
}
static bool tma_source_ok(const float *p, int X, int Y, int Z) {
    // the i1 - 1 addressing needs every axis >= 2
    return !brick_disabled() && tma_planar_ok(p, X, Y, Z) && X >= 2 && Y >= 2 && Z >= 4;
}

bool brick_eligible(const float *src, const float *own, const float *out, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, unsigned flags) {
    (void)own; (void)out; (void)X; (void)Y; (void)Z;
    if (flags & (DFM_FIELD_IN_CL | DFM_FIELD_OUT_CL)) return false;
    return tma_source_ok(src, Xs, Ys, Zs);
}

template <int TX, int BX, int BY, int BZ>
static int launch_ss_brick_t(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                             int Y, int Z, float scale, const float *bound, float bscale, cudaStream_t st) {
    CUtensorMap tmap;
    if (!encode_map(&tmap, src, B * 3, Xs, Ys, Zs, BX, BY, BZ, 3)) return DFM_EUNSUPPORTED;
    constexpr size_t smem = 3ull * BX * BY * BZ * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_ss_brick<TX, BX, BY, BZ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_ss_brick<TX, BX, BY, BZ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_ss_brick smem attribute: %s", cudaGetErrorString(e));
        if (const char *c = getenv("DFM_SS_CARVEOUT")) {                   // tuning aid: shared-memory carve-out, percent
            cudaFuncSetAttribute(k_ss_brick<TX, BX, BY, BZ, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
            cudaFuncSetAttribute(k_ss_brick<TX, BX, BY, BZ, true>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c));
        }
        configured = true;
    }
    const int nzt = (Z + TZ - 1) / TZ, nyt = (Y + TY - 1) / TY, nxt = (X + TX - 1) / TX;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    if (scale == 1.f)
        k_ss_brick<TX, BX, BY, BZ, false><<<grid, block, smem, st>>>(tmap, src, own, out, Xs, Ys, Zs, X, Y, Z, scale, make_fastdiv(nzt), bound, bscale);
    else
        k_ss_brick<TX, BX, BY, BZ, true><<<grid, block, smem, st>>>(tmap, src, own, out, Xs, Ys, Zs, X, Y, Z, scale, make_fastdiv(nzt), bound, bscale);
    return check_launch("k_ss_brick");
}

int launch_ss_brick(const float *src, const float *own, float *out, int B, int Xs, int Ys, int Zs, int X,
                    int Y, int Z, float scale, int large_box, const float *bound, float bscale, cudaStream_t st) {
    if (src != own) bound = nullptr;                  // the bound describes `own`, the box is cut from `src`
    // x/y extent = 8 (tile) + 1 (upper corner) + 1 (straddle) + deformation slack;
    // z pitch 64 = 32 + 1 + 1 + 3 (origin alignment) + slack, and bank-conflict free
#define DFM_SS(TXv, SX, SY, SZ, LX, LY, LZ)                                                                  \
    return large_box ? launch_ss_brick_t<TXv, LX, LY, LZ>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, bound, bscale, st) \
                     : launch_ss_brick_t<TXv, SX, SY, SZ>(src, own, out, B, Xs, Ys, Zs, X, Y, Z, scale, bound, bscale, st)
    DFM_SS(4, 6, 12, 40, 8, 12, 48);     // measured best of six tile / box shapes (round 1, profiles/README.md)
#undef DFM_SS
}

// first SS step from a channels-last svf (optimistic static brick); DFM_EUNSUPPORTED if not applicable
int launch_ss_first_cl(const float *svf, float *out, int B, int X, int Y, int Z, float scale, float *absmax,
                       cudaStream_t st) {
    static const bool off = getenv("DFM_NO_FIRST_CL") != nullptr;      // tuning aid
    constexpr int TX = 4, BX = 6, BY = 12, BZ = 40;
    if (off || brick_disabled() || X < 2 || Y < 2 || Z < 4 || !tma_planar_ok(svf, X, Y, 3 * Z) || Z % 4 != 0)
        return DFM_EUNSUPPORTED;
    CUtensorMap tmap;
    if (!encode_map(&tmap, svf, B, X, Y, 3 * Z, BX, BY, 3 * BZ, 1)) return DFM_EUNSUPPORTED;
    constexpr size_t smem = 3ull * BX * BY * BZ * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_ss_first_cl<TX, BX, BY, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_ss_first_cl smem attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int nzt = (Z + TZ - 1) / TZ, nyt = (Y + TY - 1) / TY, nxt = (X + TX - 1) / TX;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    k_ss_first_cl<TX, BX, BY, BZ><<<grid, block, smem, st>>>(tmap, svf, out, X, Y, Z, scale, make_fastdiv(nzt), absmax);
    return check_launch("k_ss_first_cl");
}

template <int TX, int BX, int BY, int BZ>
static int launch_warp_brick_t(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X,
                               int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st) {
    CUtensorMap tmap;
    if (!encode_map(&tmap, img, B, Xi, Yi, Zi, BX, BY, BZ, 1)) return DFM_EUNSUPPORTED;
    constexpr size_t smem = (size_t)BX * BY * BZ * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_brick smem attribute: %s", cudaGetErrorString(e));
        // Tiles whose box does not fit gather through L1.  Left alone, the driver sizes the carve-out for
        // the 7 CTAs/SM the 30 KB brick allows and leaves almost no L1 (measured 1.45 ms on the bench
        // field); 80 % (6 CTAs/SM, ~34 KB of L1) runs the same launch in 0.93 ms.
        const int carve = getenv("DFM_WARP_CARVEOUT") ? atoi(getenv("DFM_WARP_CARVEOUT")) : 80;
        cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 1, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        configured = true;
    }
    const int nzt = (Z + TZ - 1) / TZ, nyt = (Y + TY - 1) / TY, nxt = (X + TX - 1) / TX;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    const FastDiv nz = make_fastdiv(nzt);
    UpsampleArgs none = {};
#define DFM_WB(FM, HFv) k_warp_brick<TX, BX, BY, BZ, FM, HFv><<<grid, block, smem, st>>>(tmap, tmap, img, field, out, Xi, Yi, Zi, X, Y, Z, fill, nz, none)
    if (flags & DFM_FIELD_IN_CL) {
        if (has_fill) DFM_WB(1, true); else DFM_WB(1, false);
    } else {
        if (has_fill) DFM_WB(0, true); else DFM_WB(0, false);
    }
#undef DFM_WB
    return check_launch("k_warp_brick");
}

// fused RescaleTransform(zoom >= 1) + one-channel linear warp
template <int TX, int BX, int BY, int BZ>
static int launch_rescale_warp_t(const float *img, const float *half, float *out, const float *cx, const float *cy,
                                 const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y,
                                 int Z, float pre, int has_fill, float fill, cudaStream_t st) {
    CUtensorMap tmap;
    if (!encode_map(&tmap, img, B, Xi, Yi, Zi, BX, BY, BZ, 1)) return DFM_EUNSUPPORTED;
    auto ext = [](int tile, int n_in, int n_out) {
        const double ratio = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
        return (int)(tile * ratio) + 3;
    };
    const int nzt = (Z + TZ - 1) / TZ, nyt = (Y + TY - 1) / TY, nxt = (X + TX - 1) / TX;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    const FastDiv nz = make_fastdiv(nzt);
#if !DFM_EXACT_ORDER
    // default build: separable evaluation from a TMA-staged coarse box (FMODE 3) when the box covers the tile
    static const bool no_sep = getenv("DFM_NO_FUSED_SEP") != nullptr;                    // tuning aid
    if (!no_sep && ext(TX, Xh, X) <= HBX && ext(TY, Yh, Y) <= HBY && ext(TZ, Zh, Z) + 3 <= HBZ &&
        tma_source_ok(half, Xh, Yh, Zh)) {
        CUtensorMap tmap_h;
        if (!encode_map(&tmap_h, half, B * 3, Xh, Yh, Zh, HBX, HBY, HBZ, 3)) return DFM_EUNSUPPORTED;
        UpsampleArgs up = {cx, cy, cz, Xh, Yh, Zh, pre, 0};
        constexpr size_t smem = ((size_t)BX * BY * BZ + 3 * HCS) * sizeof(float);
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_brick(fused) smem attribute: %s", cudaGetErrorString(e));
            const int carve = getenv("DFM_WARP_CARVEOUT") ? atoi(getenv("DFM_WARP_CARVEOUT")) : 80;   // see launch_warp_brick_t
            cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 3, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 3, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            configured = true;
        }
        if (has_fill)
            k_warp_brick<TX, BX, BY, BZ, 3, true><<<grid, block, smem, st>>>(tmap, tmap_h, img, half, out, Xi, Yi, Zi, X, Y, Z, fill, nz, up);
        else
            k_warp_brick<TX, BX, BY, BZ, 3, false><<<grid, block, smem, st>>>(tmap, tmap_h, img, half, out, Xi, Yi, Zi, X, Y, Z, fill, nz, up);
        return check_launch("k_warp_brick(fused rescale, separable)");
    }
#endif
    UpsampleArgs up = {cx, cy, cz, Xh, Yh, Zh, pre, ext(TX, Xh, X) * ext(TY, Yh, Y) * ext(TZ, Zh, Z)};
    const size_t smem = (size_t)BX * BY * BZ * sizeof(float) + (size_t)up.cap * sizeof(float4);
    if (smem > 200 * 1024) return DFM_EUNSUPPORTED;
    static size_t configured2 = 0;
    if (smem > configured2) {
        cudaError_t e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_warp_brick<TX, BX, BY, BZ, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_brick(fused) smem attribute: %s", cudaGetErrorString(e));
        configured2 = smem;
    }
    if (has_fill)
        k_warp_brick<TX, BX, BY, BZ, 2, true><<<grid, block, smem, st>>>(tmap, tmap, img, half, out, Xi, Yi, Zi, X, Y, Z, fill, nz, up);
    else
        k_warp_brick<TX, BX, BY, BZ, 2, false><<<grid, block, smem, st>>>(tmap, tmap, img, half, out, Xi, Yi, Zi, X, Y, Z, fill, nz, up);
    return check_launch("k_warp_brick(fused rescale)");
}

int launch_rescale_warp(const float *img, const float *half, float *out, const float *cx, const float *cy,
                        const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                        float pre, int has_fill, float fill, cudaStream_t st) {
    if (!tma_source_ok(img, Xi, Yi, Zi) || Xh < 2 || Yh < 2 || Zh < 2) return DFM_EUNSUPPORTED;
    return launch_rescale_warp_t<4, 10, 16, 48>(img, half, out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, pre, has_fill, fill, st);
}

static int launch_warp_brick_var(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X,
                                 int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st) {
    WarpMaps maps;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j)
            if (!encode_map(&maps.m[i][j], img, B, Xi, Yi, Zi, 1, 12 + 4 * i, 24 + 8 * j, 1)) return DFM_EUNSUPPORTED;
    // Brick bytes per CTA.  The kernel is latency-bound (field load -> box reduction -> TMA -> gather is one serial
    // chain per tile; ncu: long-scoreboard stalls 7.5 per issue), so resident CTAs count for more than the last
    // few percent of fit rate: measured on the bench field at B=32 (carve-out 80 %): 20 KB 0.92 ms, 24 KB 0.85 ms,
    // 28 KB 0.81 ms, 32 KB 0.86 ms, 36 KB 0.85 ms, 40 KB 0.99 ms (k_warp_brick, fixed 30 KB box: 0.93 ms).
    static const int cap_kb = getenv("DFM_WARP_CAP_KB") ? atoi(getenv("DFM_WARP_CAP_KB")) : 28;      // tuning aid
    const size_t smem = (size_t)cap_kb * 1024;
    static bool configured = false;
    const int carve = getenv("DFM_WARP_CARVEOUT") ? atoi(getenv("DFM_WARP_CARVEOUT")) : 80;       // leaves L1 for the field / output streams
#define DFM_CFGV(FC, HFv)                                                                                                   \
    {                                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(k_warp_brick_var<FC, HFv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_warp_brick_var smem attribute: %s", cudaGetErrorString(e));             \
        cudaFuncSetAttribute(k_warp_brick_var<FC, HFv>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);             \
    }
    if (!configured) {
        DFM_CFGV(false, false) DFM_CFGV(false, true) DFM_CFGV(true, false) DFM_CFGV(true, true)
        configured = true;
    }
#undef DFM_CFGV
    const int nzt = (Z + 15) / 16, nyt = (Y + 7) / 8, nxt = (X + 7) / 8;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    const FastDiv nz = make_fastdiv(nzt);
    const int cap = (int)(smem / sizeof(float));
#define DFM_WBV(FC, HFv) k_warp_brick_var<FC, HFv><<<grid, block, smem, st>>>(maps, img, field, out, Xi, Yi, Zi, X, Y, Z, fill, nz, cap)
    if (flags & DFM_FIELD_IN_CL) {
        if (has_fill) DFM_WBV(true, true); else DFM_WBV(true, false);
    } else {
        if (has_fill) DFM_WBV(false, true); else DFM_WBV(false, false);
    }
#undef DFM_WBV
    return check_launch("k_warp_brick_var");
}

int launch_warp_brick(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X,
                      int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st) {
    static const bool direct = getenv("DFM_WARP_DIRECT") != nullptr;   // tuning aid
    if (direct || (flags & DFM_LOC_ABSOLUTE)) return DFM_EUNSUPPORTED;
    if (!tma_source_ok(img, Xi, Yi, Zi)) return DFM_EUNSUPPORTED;
    static const bool fixed_box = getenv("DFM_WARP_FIXED_BOX") != nullptr;      // tuning aid: the round-1 fixed-box kernel
    if (fixed_box) return launch_warp_brick_t<4, 10, 16, 48>(img, field, out, B, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
    return launch_warp_brick_var(img, field, out, B, Xi, Yi, Zi, X, Y, Z, has_fill, fill, flags, st);
}

}  // namespace dfm
