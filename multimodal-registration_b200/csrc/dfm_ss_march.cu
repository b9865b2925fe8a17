// Scaling-and-squaring step as a PLANE MARCH through a TMA ring of full-z rows (planar or
// channels-last input, planar output):   out = v + scale * interp(src, p + v),  v = scale * src.
//
// Why: the bounding-box brick (dfm_brick.cu) is shared-memory bound -- its 40-float row pitch makes
// every warp whose 32 lanes (consecutive z) straddle two x or y corner rows pay a 2-way bank conflict
// (47 % of the shared wavefronts in round 1).  A conflict-free brick needs a row pitch that is a
// multiple of 32 banks, i.e. rows as long as the volume's z extent (96 = 3 x 32 at half resolution):
// bank = z mod 32 whatever the row.  With full-z rows the z halo disappears, and the x/y halo only
// has to cover the displacement, not the tile's slant.
//
// Structure (one persistent CTA = TY y-rows x all z, marching along x):
//  * warp-specialised: TY/VPT*NZW consumer warps (warp = VPT y rows x 32 z, lane = z) + 1 producer warp.
//  * the producer streams x planes {rows y0-H .. y0+TY-1+H, all z, 3 components} through an R-slot
//    ring, one TMA box per plane (NZW boxes of 32 z x 3 for a channels-last source), `full` mbarrier
//    per slot (expect_tx), `empty` mbarrier per slot (one arrival per consumer warp).
//  * a consumer thread owns VPT (y, z) columns and marches x: own vectors from the ring's centre plane
//    (no global loads at all besides the TMA), 24 corner LDS with immediate offsets, stores planar.  In the
//    default build the two voxels of a thread share packed f32x2 maths (FADD2 / FMUL2 / FFMA2).
//  * static halo H: a voxel whose corners lie within +-H planes / rows is served by the ring.  The test
//    is on the integer corner indices, per warp (`__all_sync`); a warp with an outlier lane gathers
//    from global memory instead, so results never depend on H (bit-identical to the other kernels).
//    On the bench field (std-3 SVF) H = 2 serves 99.7 % of the warps up to the second-last step
//    and H = 3 96.9 % of the last step -- the field's maximum is far above its typical magnitude.
//  * L2 -> SM amplification (TY + 2H)/TY (+ 2H planes per ~86-step range) instead of the brick's 2.8-3.75.
//
// Measured (B200, B=32 x 80x80x96, ncu in profiles/r2_*): 116-118 us per halo-2 step (was 144 us with the
// bounding-box brick), 62 % of the HBM roofline.  The kernel is bound by the SHARED-MEMORY DATA PIPE:
// 27 loads + 3 stores per voxel are 35 wavefronts per 32 voxels (84 % of the pipe's peak in the profile);
// the remaining bank conflicts are the pigeonhole kind -- a warp whose 32 lanes span 33 columns because the
// field stretches along z.  Experiments that isolate one pipe (scripts/exp/exp_march.cu): without the upper x
// plane's 12 loads 103 us, without any corner load 87 us (the streaming floor of the ring), with trivial
// weights 115 us -- time follows the shared-memory loads, not the instruction count.
//
// Reference semantics: vxm.utils.integrate_vec, one squaring step.
#include <cuda.h>
#include <stdlib.h>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

#ifndef MARCH_EXP
#define MARCH_EXP 0          // scripts/exp/exp_march.cu: deliberately wrong variants that isolate one pipe
#endif

namespace dfm {

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// sel_mode: 0 = always run; 1 = run iff sel[b]*sel_scale <  sel_thr; 2 = run iff !(sel[b]*sel_scale < sel_thr)
struct MarchSel {
    const float *sel;
    float scale, thr;
    int mode;
};

// mbarrier helpers on 32-bit shared addresses (kept in registers and advanced by 8 per slot)
__device__ __forceinline__ bool mbar_try_wait_u(uint32_t bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow_u(uint32_t bar, uint32_t phase) {
    // bounded spin: a barrier that never completes traps instead of hanging the GPU
    for (uint32_t it = 0; !mbar_try_wait_u(bar, phase); ++it)
        if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t phase) {
    if (!mbar_try_wait_u(bar, phase)) mbar_wait_slow_u(bar, phase);
}
__device__ __forceinline__ void mbar_arrive_u(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// CTAs per SM the ring allows (227 KB of shared memory per SM, 1 KB reserved per CTA)
constexpr int march_ctas(int TY, int H, int R, int NZW) {
    return (R * 3 * (TY + 2 * H) * NZW * 32 * 4 + 2048) * 2 <= 227 * 1024 ? 2 : 1;
}

// Work decomposition: the B * nstrips * X plane-steps of a launch form one sequence (item = (batch item,
// y strip), then x); CTA c of a PERSISTENT grid owns the contiguous range [c T/G, (c+1) T/G).  Inside a range
// the ring streams without interruption -- also across the boundary between two strips, where the producer
// has the next strip's first planes in flight while the consumers finish the previous one -- so the pipeline
// fills once per CTA instead of once per tile, every CTA does the same number of steps (no wave tail), and
// the x halo is re-read once per range (~86 steps) instead of once per segment.
struct MarchRun {
    int b, y0, xs, xe, p_first, p_last;
};

// TY rows per CTA, VPT of them per thread (rows ty, ty + TY/VPT, ...): the per-step overhead (ring
// bookkeeping, barrier wait / arrive, loop) is shared by VPT voxels and their gathers interleave.
// IN_CL: 0 = planar source; 1 = channels-last source in 96-float sub-boxes (any Z); 2 = channels-last source as contiguous rows of
// 3 Z floats through one 5-D box per plane (Z a multiple of 32): corner addressing with immediate offsets like the planar path
// MEAS: max |out| per batch item goes to absmax (the first step always measures; the two steps before the last measure so that the
// per-item halo choice of the last two steps rests on the displacements actually present, not on a bound doubled per step)
template <int TY, int VPT, int H, int R, int NZW, int IN_CL, bool FIRST, bool MEAS>
__global__ void __launch_bounds__((TY / VPT * NZW + 1) * 32, march_ctas(TY, H, R, NZW))
k_ss_march(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src, float *__restrict__ out, int X,
           int Y, int Z, float scale, int nstrips, unsigned long long total, float *__restrict__ absmax, MarchSel sel) {
    constexpr int ZP = NZW * 32, ROWS = TY + 2 * H, TYW = TY / VPT, NCW = TYW * NZW;
    constexpr int CS = ROWS * ZP;          // planar: component stride in a slot; channels-last: sub-box (32 z x 3) stride / 3
    constexpr int SLOT = 3 * CS;           // floats per ring slot (one x plane)
    static_assert(R >= 2 * H + 2, "ring too short for the halo");
    static_assert(TY % VPT == 0, "rows per thread must divide the strip");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *ring = reinterpret_cast<float *>(smem_raw);
    __shared__ __align__(8) uint64_t full[R], empty[R];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t N = (uint32_t)X * Y * Z;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], NCW);
        }
    }
    __syncthreads();
    const uint32_t full_u = smem_u32(&full[0]), empty_u = smem_u32(&empty[0]);

    // Per-item variant selection (sel.mode): the batch items this launch serves are COMPACTED before the plane-steps are dealt
    // out, so that a batch split between the two halo variants keeps every CTA of both launches busy (dealing out all items
    // and skipping the foreign ones left whole CTAs idle: a mixed batch cost up to 15 % of a step).  Up to 64 items: every warp
    // reads the per-item figures once (lane = item, one ballot per 32 items) and keeps the selection as a 64-bit mask;
    // larger batches skip as before.
    auto selected = [&](int b) -> bool {
        const bool below = __ldg(sel.sel + b) * sel.scale < sel.thr;   // false for NaN
        return (sel.mode == 1) == below;
    };
    const uint32_t per_item = (uint32_t)nstrips * (uint32_t)X;
    const int nb = (int)(total / per_item);
    const bool compact = sel.mode != 0 && nb <= 64;
    unsigned long long mine = total;
    uint32_t mask_lo = 0, mask_hi = 0;
    if (compact) {                                                     // reached by every lane of every warp (before the role split)
        mask_lo = __ballot_sync(0xffffffffu, lane < nb && selected(lane));
        if (nb > 32) mask_hi = __ballot_sync(0xffffffffu, 32 + lane < nb && selected(32 + lane));
        mine = (unsigned long long)(__popc(mask_lo) + __popc(mask_hi)) * per_item;
    }
    uint32_t t = (uint32_t)(mine * blockIdx.x / gridDim.x);            // total < 2^32 (host)
    const uint32_t t1 = (uint32_t)(mine * (blockIdx.x + 1) / gridDim.x);
    // next run of this CTA's range (identical in every thread): false when the range is exhausted
    auto next_run = [&](MarchRun &r) -> bool {
        while (t < t1) {
            const uint32_t item = t / (uint32_t)X;
            r.xs = (int)(t - item * (uint32_t)X);
            r.b = (int)(item / (uint32_t)nstrips);
            r.y0 = (int)(item - (uint32_t)r.b * (uint32_t)nstrips) * TY;
            r.xe = (int)min((uint32_t)X, (uint32_t)r.xs + (t1 - t));
            t += (unsigned)(r.xe - r.xs);
            if (compact) {                                             // r.b counts selected items: map it to the batch index
                const int nlo = __popc(mask_lo);
                r.b = r.b < nlo ? (int)__fns(mask_lo, 0, r.b + 1) : 32 + (int)__fns(mask_hi, 0, r.b - nlo + 1);
            } else if (sel.mode && !selected(r.b)) {
                continue;
            }
            r.p_first = max(r.xs - H, 0);
            r.p_last = min(r.xe - 1 + H, X - 1);
            return true;
        }
        return false;
    };

    if (warp == NCW) {
        // ------------------------------ producer ------------------------------------------
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1, q = 0;                                    // ph: parity of the previous use of the slot
            MarchRun r;
            while (next_run(r)) {
                for (int p = r.p_first; p <= r.p_last; ++p, ++q) {
                    if (q >= R) mbar_wait_u(empty_u + 8 * s, ph);
                    mbar_expect_tx(&full[s], (uint32_t)(SLOT * sizeof(float)));
                    float *dst = ring + s * SLOT;
                    if (IN_CL == 2) {
                        tma_load_5d(dst, &tmap, &full[s], 0, 0, r.y0 - H, p, r.b);
                    } else if (IN_CL == 1) {
#pragma unroll
                        for (int j = 0; j < NZW; ++j) tma_load_4d(dst + j * (ROWS * 96), &tmap, &full[s], 96 * j, r.y0 - H, p, r.b);
                    } else {
                        tma_load_4d(dst, &tmap, &full[s], 0, r.y0 - H, p, r.b * 3);
                    }
                    if (++s == R) { s = 0; ph ^= 1u; }
                }
            }
        }
        return;
    }
    // ------------------------------ consumers -----------------------------------------
    const int tyw = warp / NZW, zw = warp - tyw * NZW;
    const int z = zw * 32 + lane;
    const int zc = min(z, Z - 1);                                      // threads past the edge shadow the last voxel
    const int mxi = X - 1, myi = Y - 1, mzi = Z - 1;
    const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
    const float fz = (float)zc;
    const uint32_t XS = (uint32_t)Y * Z, GX = XS, GY = (uint32_t)Z;
    const float *const ring_end = ring + R * SLOT;
    uint32_t wa = full_u, wph = 0;                                     // next plane (in load order) to wait for
    int wq = 0, rq = 0, rslot = 0;                                     // planes waited for / loaded before this run; slot of the run's first plane
    MarchRun r;
    while (next_run(r)) {
        const int rb1 = r.y0 - H + 1;                                  // ring row k holds volume row y0 - H + k
        float fy[VPT];
        int own_off[VPT];
        float *op[VPT];                                                // output pointer of component 0 at plane x
        bool ok[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int y = r.y0 + tyw + j * TYW, yc = min(y, Y - 1);
            ok[j] = (z < Z) && (y < Y);
            fy[j] = (float)yc;
            own_off[j] = IN_CL == 2 ? ((yc - (r.y0 - H)) * (3 * ZP) + 3 * zc)
                         : IN_CL == 1 ? ((zc >> 5) * (ROWS * 96) + (yc - (r.y0 - H)) * 96 + 3 * (zc & 31)) : ((yc - (r.y0 - H)) * ZP + zc);
            op[j] = out + (size_t)r.b * 3 * N + ((uint32_t)r.xs * XS + (uint32_t)yc * Z + zc);
        }
        int sb = rslot + (r.xs - H - r.p_first);                       // slot of plane xs - H (in [-H, 0] relative to the run)
        if (sb < 0) sb += R;
        int sc = sb + H;
        if (sc >= R) sc -= R;
        const float *pl = ring + sb * SLOT;                            // slot of plane x - H
        const float *pc = ring + sc * SLOT;                            // slot of plane x
        uint32_t ea = empty_u + 8 * sb;
        const int rel_x = r.p_first + H;                               // first x whose plane x - H belongs to the run
        int xmh1 = r.xs - H + 1;
        float fx = (float)r.xs;
        int am = 0;                                                    // max |out| as float bits (NaN orders above +inf)
        float amf = 0.f;

        for (int x = r.xs; x < r.xe; ++x) {
            const int need = rq + min(x + H, r.p_last) - r.p_first;
            while (wq <= need) {
                mbar_wait_u(wa, wph);
                wa += 8; ++wq;
                if (wa == full_u + 8 * R) { wa = full_u; wph ^= 1u; }
            }
#if !DFM_EXACT_ORDER
            // default build, two rows per thread: the sampling maths and the 8-term accumulations of the two
            // voxels run on the packed f32x2 pipe (FADD2 / FMUL2 / FFMA2: same roundings as the scalar fused
            // path, so the result does not depend on which path a voxel takes)
            bool packed_done = false;
            if (VPT == 2) {
                const float *poA = pc + own_off[0], *poB = pc + own_off[1];
                float vA0, vA1, vA2, vB0, vB1, vB2;
                if (IN_CL) { vA0 = poA[0]; vA1 = poA[1]; vA2 = poA[2]; vB0 = poB[0]; vB1 = poB[1]; vB2 = poB[2]; }
                else { vA0 = poA[0]; vA1 = poA[CS]; vA2 = poA[2 * CS]; vB0 = poB[0]; vB1 = poB[CS]; vB2 = poB[2 * CS]; }
                u64_t v0 = pk(vA0, vB0), v1 = pk(vA1, vB1), v2 = pk(vA2, vB2);
                const u64_t sc2 = pk(scale, scale);
                if (FIRST) { v0 = mul2(sc2, v0); v1 = mul2(sc2, v1); v2 = mul2(sc2, v2); }
                float lxA, lxB, lyA, lyB, lzA, lzB;
                upk(add2(pk(fx, fx), v0), lxA, lxB);
                upk(add2(pk(fy[0], fy[1]), v1), lyA, lyB);
                upk(add2(pk(fz, fz), v2), lzA, lzB);
                const float cxA = axis_clip(lxA, mxf), cxB = axis_clip(lxB, mxf), cyA = axis_clip(lyA, myf), cyB = axis_clip(lyB, myf),
                            czA = axis_clip(lzA, mzf), czB = axis_clip(lzB, mzf);
                const int ixA = axis_clipped_i1(cxA, mxi), ixB = axis_clipped_i1(cxB, mxi), iyA = axis_clipped_i1(cyA, myi),
                          iyB = axis_clipped_i1(cyB, myi), izA = axis_clipped_i1(czA, mzi), izB = axis_clipped_i1(czB, mzi);
                const int dxA = ixA - xmh1, dxB = ixB - xmh1, ryA = iyA - rb1, ryB = iyB - rb1;
                const bool inside = ((unsigned)dxA <= (unsigned)(2 * H - 1)) && ((unsigned)ryA <= (unsigned)(ROWS - 2)) &&
                                    ((unsigned)dxB <= (unsigned)(2 * H - 1)) && ((unsigned)ryB <= (unsigned)(ROWS - 2));
                if (__all_sync(0xffffffffu, inside)) {
                    packed_done = true;
                    const u64_t m1 = pk(-1.f, -1.f), p1 = pk(1.f, 1.f);
                    const u64_t x0 = fma2(pk(cxA, cxB), m1, pk((float)ixA, (float)ixB)), x1 = fma2(x0, m1, p1);   // w_lo = i1 - cl, w_hi = 1 - w_lo
                    const u64_t y0 = fma2(pk(cyA, cyB), m1, pk((float)iyA, (float)iyB)), y1 = fma2(y0, m1, p1);
                    const u64_t z0 = fma2(pk(czA, czB), m1, pk((float)izA, (float)izB)), z1 = fma2(z0, m1, p1);
                    const u64_t w00 = mul2(x0, y0), w01 = mul2(x0, y1), w10 = mul2(x1, y0), w11 = mul2(x1, y1);
#if MARCH_EXP == 3
                    const u64_t w[8] = {x0, x0, x0, x0, x0, x0, x0, x0};
#else
                    const u64_t w[8] = {mul2(w00, z0), mul2(w00, z1), mul2(w01, z0), mul2(w01, z1),
                                        mul2(w10, z0), mul2(w10, z1), mul2(w11, z0), mul2(w11, z1)};
#endif
                    const float *qA0 = pl + dxA * SLOT, *qB0 = pl + dxB * SLOT;
                    if (qA0 >= ring_end) qA0 -= R * SLOT;
                    if (qB0 >= ring_end) qB0 -= R * SLOT;
                    const float *qA1 = qA0 + SLOT, *qB1 = qB0 + SLOT;
                    if (qA1 == ring_end) qA1 = ring;
                    if (qB1 == ring_end) qB1 = ring;
                    u64_t acc[3];
                    if (IN_CL == 2) {
                        const int offA = ryA * (3 * ZP) + 3 * (izA - 1), offB = ryB * (3 * ZP) + 3 * (izB - 1);
                        qA0 += offA; qA1 += offA; qB0 += offB; qB1 += offB;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float va[8] = {qA0[c], qA0[3 + c], qA0[3 * ZP + c], qA0[3 * ZP + 3 + c],
                                                 qA1[c], qA1[3 + c], qA1[3 * ZP + c], qA1[3 * ZP + 3 + c]};
                            const float vb[8] = {qB0[c], qB0[3 + c], qB0[3 * ZP + c], qB0[3 * ZP + 3 + c],
                                                 qB1[c], qB1[3 + c], qB1[3 * ZP + c], qB1[3 * ZP + 3 + c]};
                            acc[c] = mul2(w[0], pk(va[0], vb[0]));
#pragma unroll
                            for (int k = 1; k < 8; ++k) acc[c] = fma2(w[k], pk(va[k], vb[k]), acc[c]);
                        }
                    } else if (IN_CL == 1) {
                        const int zlA = izA - 1, zlB = izB - 1;
                        const int faA = (zlA >> 5) * (ROWS * 96) + 3 * (zlA & 31) + ryA * 96, fbA = (izA >> 5) * (ROWS * 96) + 3 * (izA & 31) + ryA * 96;
                        const int faB = (zlB >> 5) * (ROWS * 96) + 3 * (zlB & 31) + ryB * 96, fbB = (izB >> 5) * (ROWS * 96) + 3 * (izB & 31) + ryB * 96;
                        const float *A0a = qA0 + faA, *A0b = qA0 + fbA, *A1a = qA1 + faA, *A1b = qA1 + fbA;
                        const float *B0a = qB0 + faB, *B0b = qB0 + fbB, *B1a = qB1 + faB, *B1b = qB1 + fbB;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float va[8] = {A0a[c], A0b[c], A0a[96 + c], A0b[96 + c], A1a[c], A1b[c], A1a[96 + c], A1b[96 + c]};
                            const float vb[8] = {B0a[c], B0b[c], B0a[96 + c], B0b[96 + c], B1a[c], B1b[c], B1a[96 + c], B1b[96 + c]};
                            acc[c] = mul2(w[0], pk(va[0], vb[0]));
#pragma unroll
                            for (int k = 1; k < 8; ++k) acc[c] = fma2(w[k], pk(va[k], vb[k]), acc[c]);
                        }
                    } else {
                        const int offA = ryA * ZP + (izA - 1), offB = ryB * ZP + (izB - 1);
                        qA0 += offA; qA1 += offA; qB0 += offB; qB1 += offB;
#if MARCH_EXP == 1
                        qA1 = qA0; qB1 = qB0;
#endif
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
#if MARCH_EXP == 2
                            const float oa = c == 0 ? vA0 : c == 1 ? vA1 : vA2, ob = c == 0 ? vB0 : c == 1 ? vB1 : vB2;
                            const float va[8] = {oa, oa, oa, oa, oa, oa, oa, oa}, vb[8] = {ob, ob, ob, ob, ob, ob, ob, ob};
#else
                            const float va[8] = {qA0[c * CS], qA0[c * CS + 1], qA0[c * CS + ZP], qA0[c * CS + ZP + 1],
                                                 qA1[c * CS], qA1[c * CS + 1], qA1[c * CS + ZP], qA1[c * CS + ZP + 1]};
                            const float vb[8] = {qB0[c * CS], qB0[c * CS + 1], qB0[c * CS + ZP], qB0[c * CS + ZP + 1],
                                                 qB1[c * CS], qB1[c * CS + 1], qB1[c * CS + ZP], qB1[c * CS + ZP + 1]};
#endif
                            acc[c] = mul2(w[0], pk(va[0], vb[0]));
#pragma unroll
                            for (int k = 1; k < 8; ++k) acc[c] = fma2(w[k], pk(va[k], vb[k]), acc[c]);
                        }
                    }
                    if (FIRST) { acc[0] = mul2(sc2, acc[0]); acc[1] = mul2(sc2, acc[1]); acc[2] = mul2(sc2, acc[2]); }
                    float rA0, rA1, rA2, rB0, rB1, rB2;
                    upk(add2(v0, acc[0]), rA0, rB0);
                    upk(add2(v1, acc[1]), rA1, rB1);
                    upk(add2(v2, acc[2]), rA2, rB2);
                    if (MEAS && FIRST)
                        am = max(max(max(am, __float_as_int(rA0) & 0x7fffffff), max(__float_as_int(rA1) & 0x7fffffff, __float_as_int(rA2) & 0x7fffffff)),
                                 max(__float_as_int(rB0) & 0x7fffffff, max(__float_as_int(rB1) & 0x7fffffff, __float_as_int(rB2) & 0x7fffffff)));
                    else if (MEAS)      // later steps only steer the halo choice: float max of |.| (one FMNMX per value; a NaN is skipped)
                        amf = fmaxf(fmaxf(fmaxf(amf, fabsf(rA0)), fmaxf(fabsf(rA1), fabsf(rA2))), fmaxf(fabsf(rB0), fmaxf(fabsf(rB1), fabsf(rB2))));
                    if (ok[0]) { float *o = op[0]; o[0] = rA0; o[N] = rA1; o[2 * (size_t)N] = rA2; }
                    if (ok[1]) { float *o = op[1]; o[0] = rB0; o[N] = rB1; o[2 * (size_t)N] = rB2; }
                    op[0] += XS; op[1] += XS;
                }
            }
            if (!packed_done)
#endif
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const float *po = pc + own_off[j];
                float v0, v1, v2;
                if (IN_CL) { v0 = po[0]; v1 = po[1]; v2 = po[2]; }
                else { v0 = po[0]; v1 = po[CS]; v2 = po[2 * CS]; }
                if (FIRST) { v0 = __fmul_rn(scale, v0); v1 = __fmul_rn(scale, v1); v2 = __fmul_rn(scale, v2); }
                const AxisF ax = axis_fast(__fadd_rn(fx, v0), mxf, mxi);
                const AxisF ay = axis_fast(__fadd_rn(fy[j], v1), myf, myi);
                const AxisF az = axis_fast(__fadd_rn(fz, v2), mzf, mzi);
                float w[8], a[3];
                tri_weights(ax, ay, az, w);
                const int dx = ax.i1 - xmh1;                           // ring plane of the lower x corner, relative to x - H
                const int ry = ay.i1 - rb1;                            // ring row of the lower y corner
                const bool inside = ((unsigned)dx <= (unsigned)(2 * H - 1)) && ((unsigned)ry <= (unsigned)(ROWS - 2));
                if (__all_sync(0xffffffffu, inside)) {
                    const float *q0 = pl + dx * SLOT;
                    if (q0 >= ring_end) q0 -= R * SLOT;
                    const float *q1 = q0 + SLOT;
                    if (q1 == ring_end) q1 = ring;
                    if (IN_CL == 2) {
                        const int off = ry * (3 * ZP) + 3 * (az.i1 - 1);
                        q0 += off; q1 += off;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float val[8] = {q0[c], q0[3 + c], q0[3 * ZP + c], q0[3 * ZP + 3 + c],
                                                  q1[c], q1[3 + c], q1[3 * ZP + c], q1[3 * ZP + 3 + c]};
                            a[c] = tri_accumulate(w, val);
                        }
                    } else if (IN_CL == 1) {
                        const int zl = az.i1 - 1, zh = az.i1;
                        const int fa = (zl >> 5) * (ROWS * 96) + 3 * (zl & 31) + ry * 96, fb = (zh >> 5) * (ROWS * 96) + 3 * (zh & 31) + ry * 96;
                        const float *q0a = q0 + fa, *q0b = q0 + fb, *q1a = q1 + fa, *q1b = q1 + fb;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float val[8] = {q0a[c], q0b[c], q0a[96 + c], q0b[96 + c], q1a[c], q1b[c], q1a[96 + c], q1b[96 + c]};
                            a[c] = tri_accumulate(w, val);
                        }
                    } else {
                        const int off = ry * ZP + (az.i1 - 1);
                        q0 += off; q1 += off;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float val[8] = {q0[c * CS], q0[c * CS + 1], q0[c * CS + ZP], q0[c * CS + ZP + 1],
                                                  q1[c * CS], q1[c * CS + 1], q1[c * CS + ZP], q1[c * CS + ZP + 1]};
                            a[c] = tri_accumulate(w, val);
                        }
                    }
                } else {
                    // some lane's corners leave the ring: the whole warp gathers from global memory (same arithmetic)
                    const uint32_t lo = (uint32_t)(ax.i1 - 1) * GX + (uint32_t)(ay.i1 - 1) * GY + (uint32_t)(az.i1 - 1);
                    const float *srcb = src + (size_t)r.b * 3 * N;
                    if (IN_CL) {
                        const float *g = srcb + 3 * (size_t)lo;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float val[8];
                            gather8(g + c, 3u * GY, 3u * GX, 3u, val);
                            a[c] = tri_accumulate(w, val);
                        }
                    } else {
                        const float *g = srcb + lo;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float val[8];
                            gather8(g + (size_t)c * N, GY, GX, 1u, val);
                            a[c] = tri_accumulate(w, val);
                        }
                    }
                }
                if (FIRST) { a[0] = __fmul_rn(scale, a[0]); a[1] = __fmul_rn(scale, a[1]); a[2] = __fmul_rn(scale, a[2]); }
                const float r0 = __fadd_rn(v0, a[0]), r1 = __fadd_rn(v1, a[1]), r2 = __fadd_rn(v2, a[2]);
                if (MEAS && FIRST)
                    am = max(max(am, __float_as_int(r0) & 0x7fffffff), max(__float_as_int(r1) & 0x7fffffff, __float_as_int(r2) & 0x7fffffff));
                else if (MEAS)
                    amf = fmaxf(fmaxf(amf, fabsf(r0)), fmaxf(fabsf(r1), fabsf(r2)));
                if (ok[j]) {
                    float *o = op[j];
                    o[0] = r0; o[N] = r1; o[2 * (size_t)N] = r2;
                }
                op[j] += XS;
            }
            // plane x - H is not needed by later steps of this warp (the stores above depend on every
            // shared load of the step, so the loads have completed when the arrival is issued)
            if (x >= rel_x && lane == 0) mbar_arrive_u(ea);
            ea += 8; pl += SLOT; pc += SLOT;
            if (pl == ring_end) { pl = ring; ea = empty_u; }
            if (pc == ring_end) pc = ring;
            ++xmh1;
            fx += 1.f;
        }
        // end of the run: release its remaining planes max(xe - H, p_first) .. p_last
        {
            int p = r.xe - H;
            while (p < r.p_first) { ++p; ea += 8; if (ea == empty_u + 8 * R) ea = empty_u; }
            __syncwarp();
            for (; p <= r.p_last; ++p) {
                if (lane == 0) mbar_arrive_u(ea);
                ea += 8;
                if (ea == empty_u + 8 * R) ea = empty_u;
            }
        }
        const int np = r.p_last - r.p_first + 1;
        rq += np;
        rslot = (rslot + np) % R;
        if (MEAS && absmax) {
            if (!FIRST) am = __float_as_int(amf);
            am = __reduce_max_sync(0xffffffffu, am);
            if (lane == 0) atomicMax(reinterpret_cast<int *>(absmax) + r.b, am);
        }
    }
}

// ------------------------------- host side -----------------------------------------------
static int march_cfg_int(const char *name, int dflt) {
    const char *c = getenv(name);
    return c ? atoi(c) : dflt;
}

template <int TY, int VPT, int H, int R, int NZW, int IN_CL, bool FIRST, bool MEAS>
static int launch_march_t(const float *src, float *out, int B, int X, int Y, int Z, float scale, float *absmax,
                          MarchSel sel, int seglen, cudaStream_t st) {
    constexpr int ROWS = TY + 2 * H, ZP = NZW * 32;
    constexpr size_t smem = (size_t)R * 3 * ROWS * ZP * sizeof(float);
    static_assert(smem <= 227 * 1024 - 256, "ring does not fit shared memory");
    CUtensorMap tmap;
    const bool enc = IN_CL == 2 ? encode_cl_rows_map(&tmap, src, B, X, Y, Z, ROWS)
                     : IN_CL == 1 ? encode_planar_map(&tmap, src, B, X, Y, 3 * Z, 1, ROWS, 96, 1)
                                  : encode_planar_map(&tmap, src, B * 3, X, Y, Z, 1, ROWS, ZP, 3);
    if (!enc) return DFM_EUNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_ss_march<TY, VPT, H, R, NZW, IN_CL, FIRST, MEAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_ss_march smem attribute: %s", cudaGetErrorString(e));
        cudaFuncSetAttribute(k_ss_march<TY, VPT, H, R, NZW, IN_CL, FIRST, MEAS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        configured = true;
    }
    // persistent grid: every resident CTA slot gets one contiguous range of plane-steps (at least `seglen` of them)
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int nstrip = (Y + TY - 1) / TY;
    const unsigned long long total = (unsigned long long)B * nstrip * X;
    if (total >= (1ull << 32)) return DFM_EUNSUPPORTED;
    const unsigned long long slots = (unsigned long long)sms * march_ctas(TY, H, R, NZW);
    const unsigned grid = (unsigned)max(1ull, min(slots, total / (unsigned)max(seglen, 1)));
    k_ss_march<TY, VPT, H, R, NZW, IN_CL, FIRST, MEAS><<<grid, (TY / VPT * NZW + 1) * 32, smem, st>>>(tmap, src, out, X, Y, Z, scale, nstrip, total, absmax, sel);
    return check_launch("k_ss_march");
}

template <int TY, int VPT, int H, int R, int NZW>
static int launch_march_modes(const float *src, float *out, int B, int X, int Y, int Z, float scale, bool in_cl,
                              bool first, float *absmax, MarchSel sel, int seglen, cudaStream_t st) {
    if (in_cl) {
        if (!first) return DFM_EUNSUPPORTED;
        static const bool no_rows = getenv("DFM_MARCH_NO_CL_ROWS") != nullptr;                       // tuning aid
        if (Z == NZW * 32 && !no_rows) return launch_march_t<TY, VPT, H, R, NZW, 2, true, true>(src, out, B, X, Y, Z, scale, absmax, sel, seglen, st);
        return launch_march_t<TY, VPT, H, R, NZW, 1, true, true>(src, out, B, X, Y, Z, scale, absmax, sel, seglen, st);
    }
    if (first) return launch_march_t<TY, VPT, H, R, NZW, 0, true, true>(src, out, B, X, Y, Z, scale, absmax, sel, seglen, st);
    return absmax ? launch_march_t<TY, VPT, H, R, NZW, 0, false, true>(src, out, B, X, Y, Z, scale, absmax, sel, seglen, st)
                  : launch_march_t<TY, VPT, H, R, NZW, 0, false, false>(src, out, B, X, Y, Z, scale, absmax, sel, seglen, st);
}

bool ss_march_eligible(const float *src, int X, int Y, int Z) {
    static const bool off = getenv("DFM_NO_MARCH") != nullptr || getenv("DFM_NO_BRICK") != nullptr;    // debugging aids
    return !off && X >= 2 && Y >= 2 && Z >= 33 && Z <= 128 && Z % 4 == 0 && aligned16(src) && tma_planar_ok(src, X, Y, Z);
}

// variant 0: halo 2; variant 1: halo 3 (both 2 CTAs/SM at Z <= 96) for the large displacements of the last steps.
// `first`: v = scale * src and max|out| per item goes to absmax (nullable); otherwise scale must be 1.
int launch_ss_march(const float *src, float *out, int B, int X, int Y, int Z, float scale, bool in_cl, bool first,
                    float *absmax, int variant, const float *sel, float sel_scale, float sel_thr, int sel_mode,
                    cudaStream_t st) {
    if (!ss_march_eligible(src, X, Y, Z)) return DFM_EUNSUPPORTED;
    if (!first && scale != 1.f) return DFM_EUNSUPPORTED;
    MarchSel ms = {sel, sel_scale, sel_thr, sel ? sel_mode : 0};
    static const int seg0 = march_cfg_int("DFM_MARCH_SEG", 16), seg1 = march_cfg_int("DFM_MARCH_SEG_B", 16);   // minimum steps per CTA
    const int nzw = (Z + 31) / 32;
    // block size <= 1024 threads: TY * NZW + 1 <= 32 warps; ring bytes = R * 3 * (TY + 2H) * 32 NZW * 4 <= 227 KB
#define DFM_MARCH(TYv, VPTv, Hv, Rv, NZWv, seg) \
    return launch_march_modes<TYv, VPTv, Hv, Rv, NZWv>(src, out, B, X, Y, Z, scale, in_cl, first, absmax, ms, seg, st)
    // Configurations measured on B200 at 80 x 80 x 96, B = 32: two rows per thread beat one (shared per-step
    // overhead, packed maths), an 8-slot ring beats 6 and 10, and halo 3 at 2 CTAs/SM beats halo 4 at 1 CTA/SM
    // (0.1248 vs 0.1283 ms per step).
    if (variant == 0) {
        switch (nzw) {
            case 2: DFM_MARCH(8, 2, 2, 8, 2, seg0);
            case 3: DFM_MARCH(8, 2, 2, 8, 3, seg0);
            default: DFM_MARCH(6, 2, 2, 8, 4, seg0);
        }
    }
    switch (nzw) {
        case 2: DFM_MARCH(8, 1, 4, 10, 2, seg1);
        case 3: DFM_MARCH(6, 2, 3, 8, 3, seg1);
        default: DFM_MARCH(6, 1, 4, 10, 4, seg1);
    }
#undef DFM_MARCH
}

}  // namespace dfm
