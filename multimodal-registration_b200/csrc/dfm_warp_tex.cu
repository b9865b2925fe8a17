// Fused RescaleTransform(zoom >= 1) + one-channel linear SpatialTransformer on TWO memory pipes:
//   * the full-resolution field is never materialised: a CTA owns 16 (y) x 32 (z) outputs and marches along x exactly
//     like the stand-alone up-sampler (k_upsample3_march, dfm_resize.cu): coarse planes arrive through a TMA ring, a
//     thread keeps the (y,z)-reduced coarse planes of its two output rows in registers and a field vector is one lerp
//     (default build; libdfm_exact.so keeps the reference's 8-term order).  That is ~6 shared loads per voxel on the
//     LSU pipe.
//   * the 8 image corners of a voxel are fetched by TWO texture-gather instructions (tld4) on a pitch-linear 2-D view
//     of the image volume (width = Z, height = X*Y, one texture object per batch item): tld4 returns the raw fp32
//     texels of a 2x2 (y,z) footprint, so the weights, the corner order and the accumulation stay the library's own
//     (bit-identical to dfm_resize_fwd + dfm_warp_fwd in both builds) while the gathers run on the TEX pipe and the
//     L1 acts as the brick: no bounding box, no fit test, no fallback path, no bank conflicts.
// Why: the TMA-brick warp is bound by the LSU data pipe (8 gathers x 2.3 wavefronts on the bench field)
// and every earlier fusion put the up-sampling on the same pipe.  Measured on B200 (B=32, 160x160x192): see DESIGN.md.
//
// Reference semantics: vxm.layers.RescaleTransform + vxm.layers.SpatialTransformer at the end of VxmDense
// (3d_reg.py:305,310; bids_*.py:311-322), SURVEY.md Appendix A.1-A.3, A.9.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <list>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

namespace dfm {

constexpr int WT_Y = 16, WT_Z = 32, WT_X = 32, WBY = 12, WBZ = 24, W_SLOTS = 4, W_NCW = 8;   // tile, coarse box, ring, consumer warps
constexpr int W_TEX_PER_LAUNCH = 32;

struct TexSet {
    cudaTextureObject_t t[W_TEX_PER_LAUNCH];
};

// Release of a ring slot by a consumer warp.  SYNCS.ARRIVE is NOT ordered behind the warp's in-flight shared loads (ptxas
// schedules it right after the last LDS, before their values are consumed; measured: one warp in ~1e5 saw a slot that the
// producer's next TMA box had already begun to overwrite).  The arrival therefore carries a data dependency on the
// loaded values: its count operand is 1 + (bits of the values & zero), zero being a kernel argument the host sets to 0.
__device__ __forceinline__ void wt_arrive_after(uint64_t *bar, uint32_t dep, uint32_t zero) {
    const uint32_t count = 1u + (dep & zero);               // zero is a kernel argument (0): opaque to the compilers
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// One axis of the sampling set-up with the corner index kept in FLOAT (it only feeds texture coordinates): the same
// values as axis_fast -- cl = clip(loc), i1 = min(trunc(cl) + 1, max), w_lo = i1 - cl -- without the float -> int -> float
// round trip.  trunc(cl) for 0 <= cl < 2^22 is (cl +rz 2^23) - 2^23, both steps exact.
struct AxisT {
    float i1, w0;
};
__device__ __forceinline__ AxisT axis_tex(float loc, float maxf) {
    const float cl = fminf(fmaxf(loc, 0.f), maxf);
    AxisT a;
    a.i1 = fminf(__fadd_rn(__fsub_rn(__fadd_rz(cl, 8388608.f), 8388608.f), 1.f), maxf);
    a.w0 = __fsub_rn(a.i1, cl);
    return a;
}

// a voxel pair between issuing its four tld4 and consuming them
struct TexPend {
    float4 loA, hiA, loB, hiB;
    float wxA, wyA, wzA, wxB, wyB, wzB;       // lower-corner weights per axis
    float *p;
    unsigned oob;
    uint32_t nA, nB;                          // nearest mode: the picked elements (raw bits)
};

__device__ __forceinline__ float tex_finish(const float4 &lo, const float4 &hi, float wx, float wy, float wz) {
    AxisF ax, ay, az;
    ax.w0 = wx; ax.w1 = __fsub_rn(1.f, wx); ax.i1 = 0;
    ay.w0 = wy; ay.w1 = __fsub_rn(1.f, wy); ay.i1 = 0;
    az.w0 = wz; az.w1 = __fsub_rn(1.f, wz); az.i1 = 0;
    float w[8];
    tri_weights(ax, ay, az, w);
    // tld4 component order for the footprint (i, j)..(i+1, j+1), i along the row (z), j across rows (y):
    //   .w = (i, j)   .z = (i+1, j)   .x = (i, j+1)   .y = (i+1, j+1)
    const float val[8] = {lo.w, lo.z, lo.x, lo.y, hi.w, hi.z, hi.x, hi.y};
    return tri_accumulate(w, val);
}

#if DFM_EXACT_ORDER
// reference order of the up-sampling (as upsample_emit in dfm_resize.cu): HI = plane buffer of the upper coarse plane,
// DB = row B starts one coarse row after row A (both warp-uniform)
template <int HI, bool DB>
__device__ __forceinline__ void exact_field(const float (&V)[2][3][2][3], const AxisF &ax, const AxisF &ayA, const AxisF &ayB,
                                            const AxisF &az, float post, float (&fA)[3], float (&fB)[3]) {
    constexpr int LO = HI ^ 1, RB = DB ? 1 : 0;
    float wA[8], wB[8];
    tri_weights(ax, ayA, az, wA);
    tri_weights(ax, ayB, az, wB);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float vA[8] = {V[LO][0][0][c], V[LO][0][1][c], V[LO][1][0][c], V[LO][1][1][c],
                             V[HI][0][0][c], V[HI][0][1][c], V[HI][1][0][c], V[HI][1][1][c]};
        const float vB[8] = {V[LO][RB][0][c], V[LO][RB][1][c], V[LO][RB + 1][0][c], V[LO][RB + 1][1][c],
                             V[HI][RB][0][c], V[HI][RB][1][c], V[HI][RB + 1][0][c], V[HI][RB + 1][1][c]};
        fA[c] = __fmul_rn(post, tri_accumulate(wA, vA));
        fB[c] = __fmul_rn(post, tri_accumulate(wB, vB));
    }
}
#endif

// 72 registers -> 3 CTAs (24 consumer warps) per SM.  Measured at B=32 (160x160x192): 0.727 ms; a software-pipelined variant
// (the gathers of plane x+1 issued before the results of plane x are consumed: 96 registers, 2 CTAs/SM) 0.93 ms, 4 CTAs/SM at
// 56 registers 0.728 ms -- resident warps, not gathers in flight per thread, hide the texture latency.  A hybrid that
// fetched the upper x plane's four corners with LDG (L1 hits) to relieve the texture write-back: 0.85 ms (slower).  The
// stand-alone up-sampler's __syncthreads() per coarse plane instead of the producer warp and per-warp releases: 0.80 ms
// (with gathers in flight a block barrier makes every warp wait for the slowest one's texture latency).
// NEAREST: the same march with a nearest-neighbour pick of a 4-byte image element (label maps: Transform(interp_method='nearest',
// rescale=2), 3d_reg.py:377-380) -- one plain load per voxel, values moved as raw bits, no texture involved.  Measured (B=32):
// 0.67 ms against 0.39 + 0.47 ms for the up-sampler and the stand-alone nearest warp; issue-bound (61 % busy: the march and the
// rounding / clipping of two voxels per step), storing plane x after the loads of plane x+1 (0.69 ms) and 4 CTAs/SM (0.71 ms)
// did not help.
template <bool HF, bool NEAREST>
__global__ void __launch_bounds__((W_NCW + 1) * 32, 3)
k_rescale_warp_tex(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TexSet texs, const uint32_t *__restrict__ img_bits, float *__restrict__ out,
                   const float *__restrict__ cx, const float *__restrict__ cy, const float *__restrict__ cz, int Xh, int Yh,
                   int Zh, int Xo, int Yo, int Zo, int Xi, int Yi, int Zi, float pre, float post, float fill, int nzt, int b0, uint32_t zero) {
    constexpr int SLOT_FLOATS = ((3 * WBY * WBZ + 31) / 32) * 32;
    __shared__ __align__(128) float ring[W_SLOTS][SLOT_FLOATS];
    __shared__ __align__(8) uint64_t full[W_SLOTS], empty[W_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int jz0 = zt * WT_Z, jy0 = yt * WT_Y, jx0 = blockIdx.y * WT_X;
    const int njx = min(WT_X, Xo - jx0);
    const uint32_t No = (uint32_t)Xo * Yo * Zo, XS = (uint32_t)Yo * Zo;
    const int hxi = Xh - 1, hyi = Yh - 1, hzi = Zh - 1;
    const float hxf = (float)hxi;
    // coarse box origin of this tile (tables are non-decreasing); z origin 16-byte aligned for TMA
    const int by0 = axis_fast_i1(__ldg(cy + jy0), (float)hyi, hyi) - 1;
    const int bz0 = (axis_fast_i1(__ldg(cz + jz0), (float)hzi, hzi) - 1) & ~3;
    const int px_first = axis_fast_i1(__ldg(cx + jx0), hxf, hxi) - 1;
    const int nplanes = axis_fast_i1(__ldg(cx + jx0 + njx - 1), hxf, hxi) - px_first + 1;
    const int vol0 = (b0 + (int)blockIdx.z) * 3;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < W_SLOTS; ++k) {
            mbar_init(&full[k], 1);
            mbar_init(&empty[k], W_NCW);
        }
    }
    __syncthreads();
    if (warp == W_NCW) {
        // ------------------------------ producer: coarse planes through the ring -----------------------------
        if (lane == 0)
            for (int k = 0; k < nplanes; ++k) {
                const int slot = k % W_SLOTS;
                if (k >= W_SLOTS) mbar_wait(&empty[slot], (uint32_t)((k / W_SLOTS - 1) & 1));
                mbar_expect_tx(&full[slot], (uint32_t)(3 * WBY * WBZ * sizeof(float)));
                tma_load_4d(&ring[slot][0], &tmap, &full[slot], bz0, by0, px_first + k, vol0);
            }
        return;
    }
    // ------------------------------ consumers ----------------------------------------------------------------
    const cudaTextureObject_t tex = NEAREST ? 0 : texs.t[blockIdx.z];
    const uint32_t *ibits = img_bits + (size_t)(b0 + blockIdx.z) * ((size_t)Xi * Yi * Zi);
    // per-thread geometry: rows jyA = jy0 + 2*warp, jyB = jyA + 1; column jz
    const int jyA = jy0 + 2 * warp, jz = jz0 + lane;
    const bool okA = jyA < Yo && jz < Zo, okB = (jyA + 1) < Yo && jz < Zo;
    const AxisF ayA = axis_fast(__ldg(cy + min(jyA, Yo - 1)), (float)hyi, hyi);
    const AxisF ayB = axis_fast(__ldg(cy + min(jyA + 1, Yo - 1)), (float)hyi, hyi);
    const AxisF az = axis_fast(__ldg(cz + min(jz, Zo - 1)), (float)hzi, hzi);
    const bool dB = ayB.i1 != ayA.i1;                       // row B starts one coarse row further (warp-uniform)
    const int off0 = (ayA.i1 - 1 - by0) * WBZ + (az.i1 - 1 - bz0);
    float *pA = out + (size_t)(b0 + blockIdx.z) * No + ((size_t)jx0 * Yo + min(jyA, Yo - 1)) * Zo + min(jz, Zo - 1);
    // image side
    const float mxf = (float)(Xi - 1), myf = (float)(Yi - 1), mzf = (float)(Zi - 1);
    const float fyA = (float)min(jyA, Yo - 1), fyB = (float)min(jyA + 1, Yo - 1), fz = (float)min(jz, Zo - 1);
    const float rowstep = (float)Yi;
    float fx = (float)jx0;

    TexPend pend;
    auto issue = [&](const float (&fA)[3], const float (&fB)[3], TexPend &q) {
        const float lxA = __fadd_rn(fx, fA[0]), lyA = __fadd_rn(fyA, fA[1]), lzA = __fadd_rn(fz, fA[2]);
        const float lxB = __fadd_rn(fx, fB[0]), lyB = __fadd_rn(fyB, fB[1]), lzB = __fadd_rn(fz, fB[2]);
        if (NEAREST) {
            // tf.round (half to even) on the unclipped location, then clip the integer (axis_nearest)
            const uint32_t oA = ((uint32_t)axis_nearest(lxA, Xi - 1) * Yi + axis_nearest(lyA, Yi - 1)) * Zi + axis_nearest(lzA, Zi - 1);
            const uint32_t oB = ((uint32_t)axis_nearest(lxB, Xi - 1) * Yi + axis_nearest(lyB, Yi - 1)) * Zi + axis_nearest(lzB, Zi - 1);
            uint32_t vA = __ldg(ibits + oA), vB = __ldg(ibits + oB);
            if (HF) {
                if (lxA < 0.f || lxA > mxf || lyA < 0.f || lyA > myf || lzA < 0.f || lzA > mzf) vA = __float_as_uint(fill);
                if (lxB < 0.f || lxB > mxf || lyB < 0.f || lyB > myf || lzB < 0.f || lzB > mzf) vB = __float_as_uint(fill);
            }
            q.nA = vA; q.nB = vB; q.p = pA;
            return;
        }
        q.oob = 0;
        if (HF) {
            if (lxA < 0.f || lxA > mxf || lyA < 0.f || lyA > myf || lzA < 0.f || lzA > mzf) q.oob |= 1u;
            if (lxB < 0.f || lxB > mxf || lyB < 0.f || lyB > myf || lzB < 0.f || lzB > mzf) q.oob |= 2u;
        }
        const AxisT axA = axis_tex(lxA, mxf), ayA2 = axis_tex(lyA, myf), azA = axis_tex(lzA, mzf);
        const AxisT axB = axis_tex(lxB, mxf), ayB2 = axis_tex(lyB, myf), azB = axis_tex(lzB, mzf);
        // texels (i1-1, i1) along z and rows (r, r+1), r = (ix1-1)*Yi + iy1-1: footprint centre at (iz1, r+1); integers
        // below 2^24 (host: Xi*Yi <= 65000), so the float arithmetic is exact
        const float uA = azA.i1, vA = fmaf(axA.i1, rowstep, __fsub_rn(ayA2.i1, rowstep));
        const float uB = azB.i1, vB = fmaf(axB.i1, rowstep, __fsub_rn(ayB2.i1, rowstep));
        q.loA = tex2Dgather<float4>(tex, uA, vA, 0);
        q.hiA = tex2Dgather<float4>(tex, uA, vA + rowstep, 0);
        q.loB = tex2Dgather<float4>(tex, uB, vB, 0);
        q.hiB = tex2Dgather<float4>(tex, uB, vB + rowstep, 0);
        q.wxA = axA.w0; q.wyA = ayA2.w0; q.wzA = azA.w0;
        q.wxB = axB.w0; q.wyB = ayB2.w0; q.wzB = azB.w0;
        q.p = pA;
    };
    auto finish = [&](const TexPend &q) {
        if (NEAREST) {
            uint32_t *po = reinterpret_cast<uint32_t *>(q.p);
            if (okA) __stcs(po, q.nA);
            if (okB) __stcs(po + Zo, q.nB);
            return;
        }
        float ra = tex_finish(q.loA, q.hiA, q.wxA, q.wyA, q.wzA);
        float rb = tex_finish(q.loB, q.hiB, q.wxB, q.wyB, q.wzB);
        if (HF) {
            if (q.oob & 1u) ra = fill;
            if (q.oob & 2u) rb = fill;
        }
        if (okA) __stcs(q.p, ra);
        if (okB) __stcs(q.p + Zo, rb);
    };

    static_assert(WT_X == 32, "one x coordinate per lane");
#if !DFM_EXACT_ORDER
    // separable evaluation (see k_upsample3_march): each coarse plane is reduced once to its (y,z)-bilinear value at
    // this thread's two output rows; a field vector is one lerp between the two held planes
    const float sc = __fmul_rn(pre, post);
    const float wA[4] = {sc * ayA.w0 * az.w0, sc * ayA.w0 * az.w1, sc * ayA.w1 * az.w0, sc * ayA.w1 * az.w1};
    const float wB[4] = {sc * ayB.w0 * az.w0, sc * ayB.w0 * az.w1, sc * ayB.w1 * az.w0, sc * ayB.w1 * az.w1};
    const int offB = off0 + (dB ? WBZ : 0);
    float lo[2][3], hi[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) lo[r][c] = hi[r][c] = 0.f;
    const AxisF ax_lane = axis_fast(__ldg(cx + min(jx0 + lane, Xo - 1)), hxf, hxi);
    int have = -1;
    for (int j = 0; j < njx; ++j, pA += XS, fx += 1.f) {
        const int i1 = __shfl_sync(0xffffffffu, ax_lane.i1, j);
        const float w0 = __shfl_sync(0xffffffffu, ax_lane.w0, j), w1 = __shfl_sync(0xffffffffu, ax_lane.w1, j);
        const int need = i1 - px_first;
        while (have < need) {
            ++have;
            const int slot = have % W_SLOTS;
            mbar_wait(&full[slot], (uint32_t)((have / W_SLOTS) & 1));
            const float *pa = &ring[slot][0] + off0, *pb = &ring[slot][0] + offB;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[0][c] = hi[0][c];
                lo[1][c] = hi[1][c];
                const float *qa = pa + c * (WBY * WBZ), *qb = pb + c * (WBY * WBZ);
                hi[0][c] = fmaf(wA[3], qa[WBZ + 1], fmaf(wA[2], qa[WBZ], fmaf(wA[1], qa[1], wA[0] * qa[0])));
                hi[1][c] = fmaf(wB[3], qb[WBZ + 1], fmaf(wB[2], qb[WBZ], fmaf(wB[1], qb[1], wB[0] * qb[0])));
            }
            __syncwarp();
            if (lane == 0)                                  // the warp is done with the slot once its values have arrived
                wt_arrive_after(&empty[slot], __float_as_uint(hi[0][0]) ^ __float_as_uint(hi[0][1]) ^ __float_as_uint(hi[0][2]) ^
                                                  __float_as_uint(hi[1][0]) ^ __float_as_uint(hi[1][1]) ^ __float_as_uint(hi[1][2]), zero);
        }
        float fA[3], fB[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            fA[c] = fmaf(w1, hi[0][c], w0 * lo[0][c]);
            fB[c] = fmaf(w1, hi[1][c], w0 * lo[1][c]);
        }
        issue(fA, fB, pend);
        finish(pend);
    }
#else
    float V[2][3][2][3];                                    // [plane buffer][row][column][component]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) V[a][r][q][c] = 0.f;
    const float cx_lane = __ldg(cx + min(jx0 + lane, Xo - 1));
    int have = -1;
    for (int j = 0; j < njx; ++j, pA += XS, fx += 1.f) {
        const AxisF ax = axis_fast(__shfl_sync(0xffffffffu, cx_lane, j), hxf, hxi);
        const int need = ax.i1 - px_first;
        while (have < need) {
            ++have;
            const int slot = have % W_SLOTS;
            mbar_wait(&full[slot], (uint32_t)((have / W_SLOTS) & 1));
            const float *pl = &ring[slot][0] + off0;
            uint32_t dep = 0;
            if (have & 1) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int c = 0; c < 3; ++c) dep ^= __float_as_uint(V[1][r][q][c] = __fmul_rn(pre, pl[c * (WBY * WBZ) + r * WBZ + q]));
            } else {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int c = 0; c < 3; ++c) dep ^= __float_as_uint(V[0][r][q][c] = __fmul_rn(pre, pl[c * (WBY * WBZ) + r * WBZ + q]));
            }
            __syncwarp();
            if (lane == 0) wt_arrive_after(&empty[slot], dep, zero);   // depends on all 18 loaded values
        }
        float fA[3], fB[3];
        if (have & 1) {
            if (dB) exact_field<1, true>(V, ax, ayA, ayB, az, post, fA, fB);
            else exact_field<1, false>(V, ax, ayA, ayB, az, post, fA, fB);
        } else {
            if (dB) exact_field<0, true>(V, ax, ayA, ayB, az, post, fA, fB);
            else exact_field<0, false>(V, ax, ayA, ayB, az, post, fA, fB);
        }
        issue(fA, fB, pend);
        finish(pend);
    }
#endif
}

// ---------------------------------------------------------------------------------------
// Stand-alone one-channel linear warp with the same texture gathers: the field streams through registers (planar or
// channels-last), a thread owns one (y, z) column of NX consecutive x planes (the upper plane of voxel x is the lower plane
// of voxel x+1: an L1 hit), lane = z.  No brick, no bounding box, no fit rate: the time does not depend on how strongly the
// field deforms.  Same arithmetic as k_warp_brick_var / launch_linear (bit-identical in both builds).
// ---------------------------------------------------------------------------------------
template <bool FIELD_CL, bool HF>
__global__ void __launch_bounds__(256)
k_warp_tex(const __grid_constant__ TexSet texs, const float *__restrict__ field, float *__restrict__ out, int Xi, int Yi, int Zi,
           int X, int Y, int Z, float fill, FastDiv nzt, FastDiv nyzt, int b0) {
    constexpr int NX = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xt = (int)fast_div(blockIdx.x, nyzt), rem = (int)blockIdx.x - xt * (int)nyzt.d;
    const int yt = (int)fast_div((uint32_t)rem, nzt), zt = rem - yt * (int)nzt.d;
    const int z = zt * 32 + lane, y = yt * 8 + warp, x0 = xt * NX;
    if (z >= Z || y >= Y) return;
    const cudaTextureObject_t tex = texs.t[blockIdx.y];
    const uint32_t N = (uint32_t)X * Y * Z, XS = (uint32_t)Y * Z;
    const size_t b = (size_t)(b0 + blockIdx.y);
    const float *fb = field + b * 3 * N;
    float *ob = out + b * N;
    const uint32_t vox0 = ((uint32_t)x0 * Y + y) * Z + z;
    const float mxf = (float)(Xi - 1), myf = (float)(Yi - 1), mzf = (float)(Zi - 1);
    const float fy = (float)y, fz = (float)z, rowstep = (float)Yi;
    float l[3][NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const uint32_t vox = vox0 + (uint32_t)min(i, X - 1 - x0) * XS;            // shadow the last valid plane
        if (FIELD_CL) {
            l[0][i] = __ldg(fb + (size_t)vox * 3); l[1][i] = __ldg(fb + (size_t)vox * 3 + 1); l[2][i] = __ldg(fb + (size_t)vox * 3 + 2);
        } else {
            l[0][i] = __ldg(fb + vox); l[1][i] = __ldg(fb + N + vox); l[2][i] = __ldg(fb + 2 * (size_t)N + vox);
        }
    }
    float4 lo[NX], hi[NX];
    AxisT ax[NX], ay[NX], az[NX];
    unsigned oob = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        const float lx = __fadd_rn((float)min(x0 + i, X - 1), l[0][i]), ly = __fadd_rn(fy, l[1][i]), lz = __fadd_rn(fz, l[2][i]);
        if (HF && (lx < 0.f || lx > mxf || ly < 0.f || ly > myf || lz < 0.f || lz > mzf)) oob |= 1u << i;
        ax[i] = axis_tex(lx, mxf); ay[i] = axis_tex(ly, myf); az[i] = axis_tex(lz, mzf);
        const float u = az[i].i1, v = fmaf(ax[i].i1, rowstep, __fsub_rn(ay[i].i1, rowstep));
        lo[i] = tex2Dgather<float4>(tex, u, v, 0);
        hi[i] = tex2Dgather<float4>(tex, u, v + rowstep, 0);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        if (x0 + i >= X) break;
        float r = tex_finish(lo[i], hi[i], ax[i].w0, ay[i].w0, az[i].w0);
        if (HF && (oob & (1u << i))) r = fill;
        ob[vox0 + i * XS] = r;
    }
}

// ------------------------------- host side -----------------------------------------------
// Texture objects are descriptors over caller memory (no copy).  Creating one costs a few microseconds of host time,
// so they are cached per (device, base pointer, width, rows); the cache is bounded (least recently used entries leave).
// A descriptor only describes an address range: it stays valid (and harmless) after the memory behind it is freed,
// and a later allocation at the same address with the same geometry is described by the same descriptor.
// An evicted descriptor may still be referenced by a kernel that has been launched but has not run yet (the library never
// synchronises), so eviction only moves it to a graveyard of the same capacity; it is destroyed when `cap` further
// evictions have happened, i.e. after at least `cap` newer descriptors were created -- more launches than a stream can
// hold pending.  A launch group's descriptors (<= 32) are the most recent entries, so they cannot leave while it is built.
namespace {
struct TexKey {
    int dev;
    const void *p;
    int w, h;
    bool operator==(const TexKey &o) const { return dev == o.dev && p == o.p && w == o.w && h == o.h; }
};
struct TexKeyHash {
    size_t operator()(const TexKey &k) const {
        return std::hash<const void *>()(k.p) ^ ((size_t)k.w * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.h << 20) ^ (size_t)k.dev;
    }
};
struct TexCache {
    std::mutex mu;
    std::list<std::pair<TexKey, cudaTextureObject_t>> lru;                  // front = most recent
    std::unordered_map<TexKey, std::list<std::pair<TexKey, cudaTextureObject_t>>::iterator, TexKeyHash> map;
    std::list<cudaTextureObject_t> graveyard;                               // evicted, not yet destroyed (front = most recent)
    size_t cap = 4096;                                                      // DFM_TEX_CACHE_CAP overrides (tests use a small cache)
};
TexCache &tex_cache() {
    static TexCache *c = [] {                                               // leaked on purpose: no destruction order issues at exit
        TexCache *t = new TexCache();
        if (const char *e = getenv("DFM_TEX_CACHE_CAP")) t->cap = (size_t)max(2 * W_TEX_PER_LAUNCH, atoi(e));
        return t;
    }();
    return *c;
}
}  // namespace

static bool tex_for_volume(const float *base, int Z, int rows, cudaTextureObject_t *outp) {
    int dev = 0;
    cudaGetDevice(&dev);
    TexCache &c = tex_cache();
    std::lock_guard<std::mutex> g(c.mu);
    const TexKey key = {dev, base, Z, rows};
    auto it = c.map.find(key);
    if (it != c.map.end()) {
        c.lru.splice(c.lru.begin(), c.lru, it->second);
        *outp = it->second->second;
        return true;
    }
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<float *>(base);
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = (size_t)Z;
    rd.res.pitch2D.height = (size_t)rows;
    rd.res.pitch2D.pitchInBytes = (size_t)Z * sizeof(float);
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    // descriptor creation is not a stream operation; allow it while another thread (or this one) captures a graph
    cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
    cudaThreadExchangeStreamCaptureMode(&mode);
    cudaTextureObject_t t = 0;
    const cudaError_t e = cudaCreateTextureObject(&t, &rd, &td, nullptr);
    cudaThreadExchangeStreamCaptureMode(&mode);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    while (c.lru.size() >= c.cap) {
        c.graveyard.push_front(c.lru.back().second);
        c.map.erase(c.lru.back().first);
        c.lru.pop_back();
    }
    while (c.graveyard.size() > c.cap) {
        cudaDestroyTextureObject(c.graveyard.back());
        c.graveyard.pop_back();
    }
    c.lru.emplace_front(key, t);
    c.map[key] = c.lru.begin();
    *outp = t;
    return true;
}

struct TexLimits {
    int align, pitch_align, max_w, max_h;
};
static const TexLimits &tex_limits() {
    static TexLimits l = {0, 0, 0, 0};
    static bool init = false;
    if (!init) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&l.align, cudaDevAttrTextureAlignment, dev) != cudaSuccess) l.align = 512;
        if (cudaDeviceGetAttribute(&l.pitch_align, cudaDevAttrTexturePitchAlignment, dev) != cudaSuccess) l.pitch_align = 32;
        if (cudaDeviceGetAttribute(&l.max_w, cudaDevAttrMaxTexture2DLinearWidth, dev) != cudaSuccess) l.max_w = 0;
        if (cudaDeviceGetAttribute(&l.max_h, cudaDevAttrMaxTexture2DLinearHeight, dev) != cudaSuccess) l.max_h = 0;
        cudaGetLastError();
        init = true;
    }
    return l;
}

// the image volume of every batch item can be described as a pitch-linear 2-D texture {Z, X*Y}
static bool tex_volume_ok(const float *img, int B, int Xi, int Yi, int Zi) {
    const TexLimits &l = tex_limits();
    const size_t item = (size_t)Xi * Yi * Zi * sizeof(float);
    return l.max_w > 0 && Zi <= l.max_w && (long long)Xi * Yi <= l.max_h && ((size_t)Zi * sizeof(float)) % (size_t)l.pitch_align == 0 &&
           reinterpret_cast<uintptr_t>(img) % (size_t)l.align == 0 && (B == 1 || item % (size_t)l.align == 0) &&
           Xi >= 2 && Yi >= 2 && Zi >= 2;
}

static bool upsample_box_ok(int Xi, int Yi, int Zi, int Xo, int Yo, int Zo) {
    auto ext = [](int tile, int n_in, int n_out) {          // coarse extent of `tile` outputs (+ corner)
        const double ratio = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
        return (int)(tile * ratio) + 3;
    };
    // every output advances by at most one coarse sample, and the tile's coarse box fits the TMA box
    return Xo >= Xi && Yo >= Yi && Zo >= Zi && Xi >= 2 && Yi >= 2 && Zi >= 4 && ext(WT_Y, Yi, Yo) <= WBY &&
           ext(WT_Z, Zi, Zo) + 3 <= WBZ;
}

// nearest = false: linear interpolation by texture gathers; nearest = true: nearest-neighbour pick of 4-byte elements
// (img / out are then raw 32-bit values, fill carries the fill bits), no textures needed
static int launch_rescale_warp_march(const float *img, const float *half, float *out, const float *cx, const float *cy,
                                     const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                                     float pre, int has_fill, float fill, bool nearest, cudaStream_t st) {
    static const bool off = getenv("DFM_NO_TEX") != nullptr || getenv("DFM_NO_MARCH") != nullptr;      // debugging aids
    if (off || !upsample_box_ok(Xh, Yh, Zh, X, Y, Z) || !tma_planar_ok(half, Xh, Yh, Zh)) return DFM_EUNSUPPORTED;
    if (!nearest && !tex_volume_ok(img, B, Xi, Yi, Zi)) return DFM_EUNSUPPORTED;
    CUtensorMap tmap;
    if (!encode_planar_map(&tmap, half, B * 3, Xh, Yh, Zh, 1, WBY, WBZ, 3)) return DFM_EUNSUPPORTED;
    const int nzt = (Z + WT_Z - 1) / WT_Z, nyt = (Y + WT_Y - 1) / WT_Y, nxt = (X + WT_X - 1) / WT_X;
    static bool configured = false;
    if (!configured) {
        // the L1 is this kernel's image brick: keep the shared-memory carve-out near what the rings need (measured: 25-60 %
        // 0.73 ms, 10 % 1.19 ms -- too small a carve-out costs resident CTAs)
        const int carve = getenv("DFM_TEX_CARVEOUT") ? atoi(getenv("DFM_TEX_CARVEOUT")) : 25;
        cudaFuncSetAttribute(k_rescale_warp_tex<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_rescale_warp_tex<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_rescale_warp_tex<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_rescale_warp_tex<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        configured = true;
    }
    const uint32_t *bits = reinterpret_cast<const uint32_t *>(img);
    for (int b0 = 0; b0 < B; b0 += W_TEX_PER_LAUNCH) {
        const int nb = min(W_TEX_PER_LAUNCH, B - b0);
        TexSet ts = {};
        if (!nearest)
            for (int i = 0; i < nb; ++i)
                if (!tex_for_volume(img + (size_t)(b0 + i) * Xi * Yi * Zi, Zi, Xi * Yi, &ts.t[i])) {
                    // nothing has been launched for this call yet only if this is the first group
                    DFM_REQUIRE(b0 == 0, DFM_ECUDA, "k_rescale_warp_tex: texture object creation failed after %d items", b0 + i);
                    return DFM_EUNSUPPORTED;
                }
        dim3 grid(nzt * nyt, nxt, nb), block((W_NCW + 1) * 32);
#define DFM_RWT(HFv, NNv) k_rescale_warp_tex<HFv, NNv><<<grid, block, 0, st>>>(tmap, ts, bits, out, cx, cy, cz, Xh, Yh, Zh, X, Y, Z, Xi, Yi, Zi, pre, 1.f, fill, nzt, b0, 0u)
        if (nearest) { if (has_fill) DFM_RWT(true, true); else DFM_RWT(false, true); }
        else         { if (has_fill) DFM_RWT(true, false); else DFM_RWT(false, false); }
#undef DFM_RWT
        const int rc = check_launch("k_rescale_warp_tex");
        if (rc) return rc;
    }
    return DFM_OK;
}

int launch_rescale_warp_tex(const float *img, const float *half, float *out, const float *cx, const float *cy,
                            const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                            float pre, int has_fill, float fill, cudaStream_t st) {
    return launch_rescale_warp_march(img, half, out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, pre, has_fill, fill, false, st);
}

int launch_rescale_warp_nearest(const void *img, const float *half, void *out, const float *cx, const float *cy,
                                const float *cz, int B, int Xi, int Yi, int Zi, int Xh, int Yh, int Zh, int X, int Y, int Z,
                                float pre, int has_fill, uint32_t fill_bits, cudaStream_t st) {
    if (Xi < 1 || Yi < 1 || Zi < 1) return DFM_EUNSUPPORTED;
    float fill;
    memcpy(&fill, &fill_bits, sizeof(fill));
    return launch_rescale_warp_march((const float *)img, half, (float *)out, cx, cy, cz, B, Xi, Yi, Zi, Xh, Yh, Zh, X, Y, Z, pre,
                                     has_fill, fill, true, st);
}

int launch_warp_tex(const float *img, const float *field, float *out, int B, int Xi, int Yi, int Zi, int X, int Y, int Z,
                    int has_fill, float fill, unsigned flags, cudaStream_t st) {
    // measured on B200 (B=32, 160x160x192): 0.788 / 0.726 / 0.707 ms on the bench field x1 / x0.3 / x0 against 0.806 / 0.756 /
    // 0.751 ms for the TMA-brick kernel, so this is the default where the image can be described as a texture;
    // DFM_WARP_TEX=0 (read per call: the tests switch it) selects the brick kernel
    const char *sel = getenv("DFM_WARP_TEX");
    const bool on = sel == nullptr || atoi(sel) != 0;
    if (!on || (flags & DFM_LOC_ABSOLUTE) || !tex_volume_ok(img, B, Xi, Yi, Zi)) return DFM_EUNSUPPORTED;
    const int nzt = (Z + 31) / 32, nyt = (Y + 7) / 8, nxt = (X + 1) / 2;
    if ((unsigned long long)nzt * nyt * nxt * ((unsigned long long)nzt * nyt) >= (1ull << 32)) return DFM_EUNSUPPORTED;   // fast_div range
    const FastDiv dz = make_fastdiv((uint32_t)nzt), dyz = make_fastdiv((uint32_t)(nzt * nyt));
    for (int b0 = 0; b0 < B; b0 += W_TEX_PER_LAUNCH) {
        const int nb = min(W_TEX_PER_LAUNCH, B - b0);
        TexSet ts = {};
        for (int i = 0; i < nb; ++i)
            if (!tex_for_volume(img + (size_t)(b0 + i) * Xi * Yi * Zi, Zi, Xi * Yi, &ts.t[i])) {
                DFM_REQUIRE(b0 == 0, DFM_ECUDA, "k_warp_tex: texture object creation failed after %d items", b0 + i);
                return DFM_EUNSUPPORTED;
            }
        dim3 grid((unsigned)(nzt * nyt * nxt), nb, 1), block(256);
#define DFM_WT(FC, HFv) k_warp_tex<FC, HFv><<<grid, block, 0, st>>>(ts, field, out, Xi, Yi, Zi, X, Y, Z, fill, dz, dyz, b0)
        if (flags & DFM_FIELD_IN_CL) {
            if (has_fill) DFM_WT(true, true); else DFM_WT(true, false);
        } else {
            if (has_fill) DFM_WT(false, true); else DFM_WT(false, false);
        }
#undef DFM_WT
        const int rc = check_launch("k_warp_tex");
        if (rc) return rc;
    }
    return DFM_OK;
}

}  // namespace dfm
