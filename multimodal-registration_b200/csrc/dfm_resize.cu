// ne.utils.resize / rescale_dense_transform: resample C channels onto a separable coordinate
// grid given by per-axis tables (tf.linspace(0, n_in-1, n_out) in the reference).
//   out[b,c,jx,jy,jz] = post * interp(pre * in[b,c], (cx[jx], cy[jy], cz[jz]))
// Direct kernel: lanes own consecutive z outputs, a thread owns ROWS consecutive rows; the
// input (1/8 of the output for a x2 upsample) stays in L1/L2.
#include <cuda.h>
#include <stdlib.h>

#include "dfm_common.cuh"
#include "dfm_tma.cuh"

namespace dfm {

template <int ROWS, int INTERP, bool IN_CL, bool OUT_CL>
__global__ void __launch_bounds__(256, 4)
k_resize(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ cx,
         const float *__restrict__ cy, const float *__restrict__ cz, int C, int Xi, int Yi, int Zi,
         int Xo, int Yo, int Zo, float pre, float post, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t yy = fast_div(p, zdiv);
    const uint32_t jz = p - yy * zdiv.d;
    const uint32_t jx = blockIdx.y;
    const uint32_t No = (uint32_t)Xo * Yo * Zo, Ni = (uint32_t)Xi * Yi * Zi;
    const float *ib = in + (size_t)blockIdx.z * C * Ni;
    float *ob = out + (size_t)blockIdx.z * C * No;
    const float lx = __ldg(cx + jx), lz = __ldg(cz + jz);
    const uint32_t YZ = (uint32_t)Yi * Zi;

    if (INTERP == DFM_LINEAR) {
        const Axis ax = axis_linear(lx, (float)(Xi - 1));
        const Axis az = axis_linear(lz, (float)(Zi - 1));
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t jy = yy * ROWS + r;
            if (jy >= (uint32_t)Yo) break;
            const Axis ay = axis_linear(__ldg(cy + jy), (float)(Yi - 1));
            const uint32_t b00 = ax.i0 * YZ + ay.i0 * Zi, b01 = ax.i0 * YZ + ay.i1 * Zi;
            const uint32_t b10 = ax.i1 * YZ + ay.i0 * Zi, b11 = ax.i1 * YZ + ay.i1 * Zi;
            const float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
            const float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
            uint32_t off[8];
            float w[8];
            off[0] = b00 + az.i0; off[1] = b00 + az.i1; off[2] = b01 + az.i0; off[3] = b01 + az.i1;
            off[4] = b10 + az.i0; off[5] = b10 + az.i1; off[6] = b11 + az.i0; off[7] = b11 + az.i1;
            w[0] = __fmul_rn(w00, az.w0); w[1] = __fmul_rn(w00, az.w1);
            w[2] = __fmul_rn(w01, az.w0); w[3] = __fmul_rn(w01, az.w1);
            w[4] = __fmul_rn(w10, az.w0); w[5] = __fmul_rn(w10, az.w1);
            w[6] = __fmul_rn(w11, az.w0); w[7] = __fmul_rn(w11, az.w1);
            const uint32_t vox = (jx * Yo + jy) * Zo + jz;
            const float *ic = ib;
            float *oc = OUT_CL ? ob + (size_t)vox * C : ob + vox;
            for (int c = 0; c < C; ++c) {
                float val[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    val[k] = __fmul_rn(pre, IN_CL ? __ldg(ic + (size_t)off[k] * C) : __ldg(ic + off[k]));
                *oc = __fmul_rn(post, tri_accumulate(w, val));
                ic += IN_CL ? 1 : Ni;
                oc += OUT_CL ? 1 : No;
            }
        }
    } else {
        const uint32_t bxz = (uint32_t)axis_nearest(lx, Xi - 1) * YZ + axis_nearest(lz, Zi - 1);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t jy = yy * ROWS + r;
            if (jy >= (uint32_t)Yo) break;
            const uint32_t off = bxz + (uint32_t)axis_nearest(__ldg(cy + jy), Yi - 1) * Zi;
            const uint32_t vox = (jx * Yo + jy) * Zo + jz;
            const float *ic = IN_CL ? ib + (size_t)off * C : ib + off;
            float *oc = OUT_CL ? ob + (size_t)vox * C : ob + vox;
            for (int c = 0; c < C; ++c) {
                *oc = __fmul_rn(post, __fmul_rn(pre, __ldg(ic)));
                ic += IN_CL ? 1 : Ni;
                oc += OUT_CL ? 1 : No;
            }
        }
    }
}

template <int INTERP>
static int launch_resize(const float *in, float *out, const float *cx, const float *cy, const float *cz, int B,
                         int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre, float post,
                         unsigned flags, cudaStream_t st) {
    constexpr int ROWS = 4;
    const uint32_t plane = (uint32_t)((Yo + ROWS - 1) / ROWS) * Zo;
    dim3 grid((plane + 255) / 256, Xo, B), block(256);
    FastDiv fd = make_fastdiv(Zo);
    const bool icl = flags & DFM_FIELD_IN_CL, ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(I, O) k_resize<ROWS, INTERP, I, O><<<grid, block, 0, st>>>( \
        in, out, cx, cy, cz, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, fd, plane)
    if (icl) { if (ocl) DFM_GO(true, true); else DFM_GO(true, false); }
    else     { if (ocl) DFM_GO(false, true); else DFM_GO(false, false); }
#undef DFM_GO
    return check_launch("dfm_resize_fwd");
}

// ---------------------------------------------------------------------------------------
// Up-sampling of a planar 3-component field through shared memory (the x2 RescaleTransform of
// VxmDense).  A CTA owns 8 x 8 x 64 outputs; the input box they sample (about 6 x 6 x 34 for
// x2) is staged once as float4 {c0, c1, c2, -} so that one 128-bit shared load fetches all
// three components of a corner: 8 LDS.128 per output voxel instead of 24 LDS.32, and because
// neighbouring outputs share input voxels (pairs of lanes hit the same address -> broadcast)
// the crossbar moves half as many wavefronts.  Same op order as the direct kernel.  If the
// box does not fit the allocation (general tables), the CTA gathers from global memory.
// ---------------------------------------------------------------------------------------
constexpr int RT_X = 8, RT_Y = 8, RT_Z = 64;

__device__ __forceinline__ int axis_i0(float loc, int maxi) {
    return (int)fminf(fmaxf(floorf(loc), 0.f), (float)maxi);
}

__global__ void __launch_bounds__(256)
k_resize3_smem(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ cx,
               const float *__restrict__ cy, const float *__restrict__ cz, int Xi, int Yi, int Zi, int Xo,
               int Yo, int Zo, float pre, float post, int nzt, int cap_x, int cap_y, int cap_z) {
    extern __shared__ float4 box[];                       // [nbx][nby][nbz]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int jx0 = blockIdx.y * RT_X, jy0 = yt * RT_Y, jz0 = zt * RT_Z;
    const int jx1 = min(jx0 + RT_X, Xo) - 1, jy1 = min(jy0 + RT_Y, Yo) - 1, jz1 = min(jz0 + RT_Z, Zo) - 1;
    const uint32_t No = (uint32_t)Xo * Yo * Zo, Ni = (uint32_t)Xi * Yi * Zi;
    const float *ib = in + (size_t)blockIdx.z * 3 * Ni;
    float *ob = out + (size_t)blockIdx.z * 3 * No;
    // input box of this tile (tables are non-decreasing): [i0(first), min(i0(last) + 1, max)]
    const int bx0 = axis_i0(__ldg(cx + jx0), Xi - 1), bx1 = min(axis_i0(__ldg(cx + jx1), Xi - 1) + 1, Xi - 1);
    const int by0 = axis_i0(__ldg(cy + jy0), Yi - 1), by1 = min(axis_i0(__ldg(cy + jy1), Yi - 1) + 1, Yi - 1);
    const int bz0 = axis_i0(__ldg(cz + jz0), Zi - 1), bz1 = min(axis_i0(__ldg(cz + jz1), Zi - 1) + 1, Zi - 1);
    const int nbx = bx1 - bx0 + 1, nby = by1 - by0 + 1, nbz = bz1 - bz0 + 1;
    const bool staged = nbx <= cap_x && nby <= cap_y && nbz <= cap_z;     // CTA-uniform
    if (staged) {
        const int total = nbx * nby * nbz;
        for (int t = threadIdx.x; t < total; t += 256) {
            const int bz = t % nbz, q = t / nbz, by = q % nby, bx = q / nby;
            const uint32_t o = ((uint32_t)(bx0 + bx) * Yi + (by0 + by)) * Zi + (bz0 + bz);
            box[t] = make_float4(__fmul_rn(pre, __ldg(ib + o)), __fmul_rn(pre, __ldg(ib + Ni + o)),
                                 __fmul_rn(pre, __ldg(ib + 2 * (size_t)Ni + o)), 0.f);
        }
    }
    __syncthreads();

    const int jy = jy0 + warp;
    if (jy >= Yo) return;
    const Axis ay = axis_linear(__ldg(cy + jy), (float)(Yi - 1));
    Axis az[2];
    bool okz[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int jz = jz0 + s * 32 + lane;
        okz[s] = jz < Zo;
        az[s] = axis_linear(__ldg(cz + min(jz, Zo - 1)), (float)(Zi - 1));
    }
#pragma unroll 2
    for (int i = 0; i < RT_X; ++i) {
        const int jx = jx0 + i;
        if (jx >= Xo) break;
        const Axis ax = axis_linear(__ldg(cx + jx), (float)(Xi - 1));
        const float w00 = __fmul_rn(ax.w0, ay.w0), w01 = __fmul_rn(ax.w0, ay.w1);
        const float w10 = __fmul_rn(ax.w1, ay.w0), w11 = __fmul_rn(ax.w1, ay.w1);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!okz[s]) continue;
            float w[8];
            w[0] = __fmul_rn(w00, az[s].w0); w[1] = __fmul_rn(w00, az[s].w1);
            w[2] = __fmul_rn(w01, az[s].w0); w[3] = __fmul_rn(w01, az[s].w1);
            w[4] = __fmul_rn(w10, az[s].w0); w[5] = __fmul_rn(w10, az[s].w1);
            w[6] = __fmul_rn(w11, az[s].w0); w[7] = __fmul_rn(w11, az[s].w1);
            float v0[8], v1[8], v2[8];
            if (staged) {
                const int r00 = ((ax.i0 - bx0) * nby + (ay.i0 - by0)) * nbz - bz0;
                const int r01 = ((ax.i0 - bx0) * nby + (ay.i1 - by0)) * nbz - bz0;
                const int r10 = ((ax.i1 - bx0) * nby + (ay.i0 - by0)) * nbz - bz0;
                const int r11 = ((ax.i1 - bx0) * nby + (ay.i1 - by0)) * nbz - bz0;
                const int o[8] = {r00 + az[s].i0, r00 + az[s].i1, r01 + az[s].i0, r01 + az[s].i1,
                                  r10 + az[s].i0, r10 + az[s].i1, r11 + az[s].i0, r11 + az[s].i1};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 t = box[o[k]];
                    v0[k] = t.x; v1[k] = t.y; v2[k] = t.z;
                }
            } else {
                const uint32_t YZ = (uint32_t)Yi * Zi;
                const uint32_t b00 = ax.i0 * YZ + ay.i0 * Zi, b01 = ax.i0 * YZ + ay.i1 * Zi;
                const uint32_t b10 = ax.i1 * YZ + ay.i0 * Zi, b11 = ax.i1 * YZ + ay.i1 * Zi;
                const uint32_t o[8] = {b00 + az[s].i0, b00 + az[s].i1, b01 + az[s].i0, b01 + az[s].i1,
                                       b10 + az[s].i0, b10 + az[s].i1, b11 + az[s].i0, b11 + az[s].i1};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v0[k] = __fmul_rn(pre, __ldg(ib + o[k]));
                    v1[k] = __fmul_rn(pre, __ldg(ib + Ni + o[k]));
                    v2[k] = __fmul_rn(pre, __ldg(ib + 2 * (size_t)Ni + o[k]));
                }
            }
            const uint32_t vox = ((uint32_t)jx * Yo + jy) * Zo + (jz0 + s * 32 + lane);
            ob[vox] = __fmul_rn(post, tri_accumulate(w, v0));
            ob[No + vox] = __fmul_rn(post, tri_accumulate(w, v1));
            ob[2 * (size_t)No + vox] = __fmul_rn(post, tri_accumulate(w, v2));
        }
    }
}

// ---------------------------------------------------------------------------------------
// Plane-marching up-sampler for planar 3-component fields (zoom >= 1 on every axis).
// A CTA owns 16 (y) x 32 (z) outputs and marches along x.  The coarse planes it needs arrive
// through a TMA ring (one 4-D box {UBZ, UBY, 1, 3} per coarse plane).  A thread owns two output
// rows and one z: it keeps the 3 rows x 2 columns x 3 components it needs of the two current
// coarse planes in REGISTERS (36 floats) and only loads a new coarse plane's 18 values when the
// march crosses into it -- about 4.5 shared loads per output voxel instead of 24.  The
// accumulation is the reference's (corner order, left-to-right weight product, pre-scaled
// values), so results are bit-identical to k_resize / k_resize3_smem.
// ---------------------------------------------------------------------------------------
constexpr int UT_Y = 16, UT_Z = 32, UT_X = 32, UBY = 12, UBZ = 24, U_SLOTS = 4;

// two outputs (rows A and B of one thread) x three components from the register-held planes;
// HI = which of the two plane buffers holds the upper coarse plane, DB = row B starts one coarse
// row after row A.  Both are warp-uniform, so the four instantiations are plain branches.
template <int HI, bool DB>
__device__ __forceinline__ void upsample_emit(const float (&V)[2][3][2][3], const AxisF &ax, const AxisF &ayA,
                                              const AxisF &ayB, const AxisF &az, float post, float *pA, uint32_t No,
                                              uint32_t Zo, bool okA, bool okB) {
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
        const float ra = tri_accumulate(wA, vA), rb = tri_accumulate(wB, vB);
        if (okA) pA[(size_t)c * No] = __fmul_rn(post, ra);
        if (okB) pA[(size_t)c * No + Zo] = __fmul_rn(post, rb);
    }
}

__global__ void __launch_bounds__(256)
k_upsample3_march(const __grid_constant__ CUtensorMap tmap, float *__restrict__ out, const float *__restrict__ cx,
                  const float *__restrict__ cy, const float *__restrict__ cz, int Xi, int Yi, int Zi, int Xo,
                  int Yo, int Zo, float pre, float post, int nzt) {
    constexpr int SLOT_FLOATS = ((3 * UBY * UBZ + 31) / 32) * 32;
    __shared__ __align__(128) float ring[U_SLOTS][SLOT_FLOATS];
    __shared__ __align__(8) uint64_t bar[U_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int zt = blockIdx.x % nzt, yt = blockIdx.x / nzt;
    const int jz0 = zt * UT_Z, jy0 = yt * UT_Y, jx0 = blockIdx.y * UT_X;
    const int njx = min(UT_X, Xo - jx0);
    const uint32_t No = (uint32_t)Xo * Yo * Zo, uZo = (uint32_t)Zo, XS = (uint32_t)Yo * Zo;
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const float mxf = (float)mxi;
    // coarse box origin of this tile (tables are non-decreasing); z origin 16-byte aligned for TMA
    const int by0 = axis_fast_i1(__ldg(cy + jy0), (float)myi, myi) - 1;
    const int bz0 = (axis_fast_i1(__ldg(cz + jz0), (float)mzi, mzi) - 1) & ~3;
    const int px_first = axis_fast_i1(__ldg(cx + jx0), mxf, mxi) - 1;
    const int nplanes = axis_fast_i1(__ldg(cx + jx0 + njx - 1), mxf, mxi) - px_first + 1;
    const int vol0 = (int)blockIdx.z * 3;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < U_SLOTS; ++k) mbar_init(&bar[k], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int k = 0; k < min(nplanes, U_SLOTS); ++k) {
            mbar_expect_tx(&bar[k], (uint32_t)(3 * UBY * UBZ * sizeof(float)));
            tma_load_4d(&ring[k][0], &tmap, &bar[k], bz0, by0, px_first + k, vol0);
        }

    // per-thread geometry: rows jyA = jy0 + 2*warp, jyB = jyA + 1; column jz
    const int jyA = jy0 + 2 * warp, jz = jz0 + lane;
    const bool okA = jyA < Yo && jz < Zo, okB = (jyA + 1) < Yo && jz < Zo;
    const AxisF ayA = axis_fast(__ldg(cy + min(jyA, Yo - 1)), (float)myi, myi);
    const AxisF ayB = axis_fast(__ldg(cy + min(jyA + 1, Yo - 1)), (float)myi, myi);
    const AxisF az = axis_fast(__ldg(cz + min(jz, Zo - 1)), (float)mzi, mzi);
    const bool dB = ayB.i1 != ayA.i1;                       // row B starts one coarse row further (warp-uniform)
    const int off0 = (ayA.i1 - 1 - by0) * UBZ + (az.i1 - 1 - bz0);
    float *pA = out + (size_t)blockIdx.z * 3 * No + ((size_t)jx0 * Yo + jyA) * Zo + jz;

#if !DFM_EXACT_ORDER
    // Fast arithmetic: the interpolation is separable, and along the march only x changes.  Each
    // coarse plane is reduced ONCE to its (y,z)-bilinear value at this thread's two output rows
    // (4 taps x 3 components x 2 rows, weights fixed for the whole march and carrying pre*post);
    // an output voxel is then one lerp between the two held planes: 2 flops per component instead
    // of 8, and 12 shared loads per output pair and coarse plane.  Differs from the reference's
    // summation order by a few ulp (libdfm_exact.so keeps the reference order).
    const float sc = __fmul_rn(pre, post);
    const float wA[4] = {sc * ayA.w0 * az.w0, sc * ayA.w0 * az.w1, sc * ayA.w1 * az.w0, sc * ayA.w1 * az.w1};
    const float wB[4] = {sc * ayB.w0 * az.w0, sc * ayB.w0 * az.w1, sc * ayB.w1 * az.w0, sc * ayB.w1 * az.w1};
    const int offB = off0 + (dB ? UBZ : 0);
    float lo[2][3], hi[2][3];                               // [row A/B][component] of the two held planes
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) lo[r][c] = hi[r][c] = 0.f;
    static_assert(UT_X == 32, "one x coordinate per lane");
    const AxisF ax_lane = axis_fast(__ldg(cx + min(jx0 + lane, Xo - 1)), mxf, mxi);
    int have = -1;
    for (int j = 0; j < njx; ++j, pA += XS) {
        const int i1 = __shfl_sync(0xffffffffu, ax_lane.i1, j);
        const float w0 = __shfl_sync(0xffffffffu, ax_lane.w0, j), w1 = __shfl_sync(0xffffffffu, ax_lane.w1, j);
        const int need = i1 - px_first;
        while (have < need) {
            ++have;
            const int slot = have % U_SLOTS;
            mbar_wait(&bar[slot], (uint32_t)((have / U_SLOTS) & 1));
            const float *pa = &ring[slot][0] + off0, *pb = &ring[slot][0] + offB;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[0][c] = hi[0][c];
                lo[1][c] = hi[1][c];
                const float *qa = pa + c * (UBY * UBZ), *qb = pb + c * (UBY * UBZ);
                hi[0][c] = fmaf(wA[3], qa[UBZ + 1], fmaf(wA[2], qa[UBZ], fmaf(wA[1], qa[1], wA[0] * qa[0])));
                hi[1][c] = fmaf(wB[3], qb[UBZ + 1], fmaf(wB[2], qb[UBZ], fmaf(wB[1], qb[1], wB[0] * qb[0])));
            }
            __syncthreads();                                // every thread has read its values out of the slot
            if (threadIdx.x == 0 && have + U_SLOTS < nplanes) {
                mbar_expect_tx(&bar[slot], (uint32_t)(3 * UBY * UBZ * sizeof(float)));
                tma_load_4d(&ring[slot][0], &tmap, &bar[slot], bz0, by0, px_first + have + U_SLOTS, vol0);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (okA) pA[(size_t)c * No] = fmaf(w1, hi[0][c], w0 * lo[0][c]);
            if (okB) pA[(size_t)c * No + uZo] = fmaf(w1, hi[1][c], w0 * lo[1][c]);
        }
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

    // the x coordinates of the chunk live in registers (one per lane)
    static_assert(UT_X == 32, "one x coordinate per lane");
    const float cx_lane = __ldg(cx + min(jx0 + lane, Xo - 1));
    int have = -1;                                          // relative index of the newest coarse plane held
    for (int j = 0; j < njx; ++j, pA += XS) {
        const AxisF ax = axis_fast(__shfl_sync(0xffffffffu, cx_lane, j), mxf, mxi);
        const int need = ax.i1 - px_first;                  // relative index of the UPPER coarse plane (uniform)
        while (have < need) {                               // march: coarse planes are consumed in order
            ++have;
            const int slot = have % U_SLOTS;
            mbar_wait(&bar[slot], (uint32_t)((have / U_SLOTS) & 1));
            const float *pl = &ring[slot][0] + off0;
            if (have & 1) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int c = 0; c < 3; ++c) V[1][r][q][c] = __fmul_rn(pre, pl[c * (UBY * UBZ) + r * UBZ + q]);
            } else {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int c = 0; c < 3; ++c) V[0][r][q][c] = __fmul_rn(pre, pl[c * (UBY * UBZ) + r * UBZ + q]);
            }
            __syncthreads();                                // every thread has copied its values out of the slot
            if (threadIdx.x == 0 && have + U_SLOTS < nplanes) {
                mbar_expect_tx(&bar[slot], (uint32_t)(3 * UBY * UBZ * sizeof(float)));
                tma_load_4d(&ring[slot][0], &tmap, &bar[slot], bz0, by0, px_first + have + U_SLOTS, vol0);
            }
        }
        if (have & 1) {
            if (dB) upsample_emit<1, true>(V, ax, ayA, ayB, az, post, pA, No, uZo, okA, okB);
            else upsample_emit<1, false>(V, ax, ayA, ayB, az, post, pA, No, uZo, okA, okB);
        } else {
            if (dB) upsample_emit<0, true>(V, ax, ayA, ayB, az, post, pA, No, uZo, okA, okB);
            else upsample_emit<0, false>(V, ax, ayA, ayB, az, post, pA, No, uZo, okA, okB);
        }
    }
#endif
}

static bool upsample_march_ok(int Xi, int Yi, int Zi, int Xo, int Yo, int Zo) {
    auto ext = [](int tile, int n_in, int n_out) {          // coarse extent of `tile` outputs (+ corner)
        const double ratio = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
        return (int)(tile * ratio) + 3;
    };
    // every output advances by at most one coarse sample, and the tile's coarse box fits the TMA box
    return Xo >= Xi && Yo >= Yi && Zo >= Zi && Xi >= 2 && Yi >= 2 && Zi >= 4 && ext(UT_Y, Yi, Yo) <= UBY &&
           ext(UT_Z, Zi, Zo) + 3 <= UBZ;
}

static int launch_upsample3_march(const float *in, float *out, const float *cx, const float *cy, const float *cz,
                                  int B, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre, float post,
                                  cudaStream_t st) {
    static const bool off = getenv("DFM_NO_BRICK") != nullptr || getenv("DFM_NO_MARCH") != nullptr;
    if (off || !upsample_march_ok(Xi, Yi, Zi, Xo, Yo, Zo) || !tma_planar_ok(in, Xi, Yi, Zi)) return DFM_EUNSUPPORTED;
    CUtensorMap tmap;
    if (!encode_planar_map(&tmap, in, B * 3, Xi, Yi, Zi, 1, UBY, UBZ, 3)) return DFM_EUNSUPPORTED;
    const int nzt = (Zo + UT_Z - 1) / UT_Z, nyt = (Yo + UT_Y - 1) / UT_Y, nxt = (Xo + UT_X - 1) / UT_X;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    k_upsample3_march<<<grid, block, 0, st>>>(tmap, out, cx, cy, cz, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, nzt);
    return check_launch("dfm_resize_fwd(march)");
}

static int launch_resize3_smem(const float *in, float *out, const float *cx, const float *cy, const float *cz,
                               int B, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre, float post,
                               cudaStream_t st) {
    // box capacity from the zoom ratio (+ slack); tiles whose box is larger gather from global
    auto cap = [](int tile, int n_in, int n_out) {
        const double ratio = n_out > 1 ? (double)(n_in - 1) / (double)(n_out - 1) : 0.0;
        return (int)(tile * ratio) + 3;
    };
    const int cap_x = cap(RT_X, Xi, Xo), cap_y = cap(RT_Y, Yi, Yo), cap_z = cap(RT_Z, Zi, Zo);
    const size_t smem = (size_t)cap_x * cap_y * cap_z * sizeof(float4);
    if (smem > 64 * 1024) return DFM_EUNSUPPORTED;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_resize3_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_resize3_smem smem attribute: %s", cudaGetErrorString(e));
        configured = 64 * 1024;
    }
    const int nzt = (Zo + RT_Z - 1) / RT_Z, nyt = (Yo + RT_Y - 1) / RT_Y, nxt = (Xo + RT_X - 1) / RT_X;
    dim3 grid(nzt * nyt, nxt, B), block(256);
    k_resize3_smem<<<grid, block, smem, st>>>(in, out, cx, cy, cz, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, nzt,
                                              cap_x, cap_y, cap_z);
    return check_launch("dfm_resize_fwd(smem)");
}

// ---------------------------------------------------------------------------------------
// adjoint (gather form, no atomics):
//   gin[i] = pre*post * sum_{kx} wx[ix][kx] sum_{ky} wy[iy][ky] sum_{kz} wz[iz][kz] gout[xlo+kx, ylo+ky, zlo+kz]
// The per-axis tap tables (first output index, tap count, weights) are computed by the host from
// the same coordinate tables the forward kernel uses (`_coords.adjoint_taps`).
// ---------------------------------------------------------------------------------------
// KZ = compile-time z tap count (row length of zw): the z taps of a row are loaded together
// (memory-level parallelism); rows shorter than KZ are zero-padded by the host tables.
// A thread owns RX consecutive input planes of one (y, z).  All tap loops have compile-time bounds
// (K = the largest tap count; the host tables are zero-padded, indices are clamped), so every load of a
// thread is independent of every other and the compiler issues them together -- the first version
// (data-dependent loop bounds, one output per thread) was bound by its three-deep dependent chain
// table -> gout -> store.
template <int K, int RX>
__global__ void __launch_bounds__(256)
k_resize_bwd(const float *__restrict__ gout, float *__restrict__ gin, const int *__restrict__ xlo,
             const float *__restrict__ xw, int kx, const int *__restrict__ ylo, const float *__restrict__ yw, int ky,
             const int *__restrict__ zlo, const float *__restrict__ zw, int kz, int Xi, int Yi, int Zi, int Xo,
             int Yo, int Zo, float s, FastDiv zdiv, uint32_t plane_items) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const uint32_t iy = fast_div(p, zdiv);
    const uint32_t iz = p - iy * zdiv.d;
    const uint32_t ix0 = blockIdx.y * RX;
    const uint32_t bc = blockIdx.z;   // b*C + c
    const size_t No = (size_t)Xo * Yo * Zo, Ni = (size_t)Xi * Yi * Zi;
    const float *gb = gout + (size_t)bc * No;
    const int y0 = __ldg(ylo + iy), z0 = __ldg(zlo + iz);
    float wy[K], wz[K];
    uint32_t oy[K], oz[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {                     // taps past the table width have weight 0 and a valid address
        wz[c] = c < kz ? __ldg(zw + iz * kz + c) : 0.f;
        oz[c] = (uint32_t)min(z0 + c, Zo - 1);
        wy[c] = c < ky ? __ldg(yw + iy * ky + c) : 0.f;
        oy[c] = (uint32_t)min(y0 + c, Yo - 1) * (uint32_t)Zo;
    }
#pragma unroll
    for (int r = 0; r < RX; ++r) {
        const uint32_t ix = ix0 + r;
        if (ix >= (uint32_t)Xi) break;
        const int x0 = __ldg(xlo + ix);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const float wx = a < kx ? __ldg(xw + ix * kx + a) : 0.f;
            const float *pl = gb + (size_t)min(x0 + a, Xo - 1) * Yo * Zo;
            float accy = 0.f;
#pragma unroll
            for (int b = 0; b < K; ++b) {
                float accz = 0.f;
#pragma unroll
                for (int c = 0; c < K; ++c) accz = fmaf(wz[c], __ldg(pl + oy[b] + oz[c]), accz);
                accy = fmaf(wy[b], accz, accy);
            }
            acc = fmaf(wx, accy, acc);
        }
        gin[(size_t)bc * Ni + ((size_t)ix * Yi + iy) * Zi + iz] = s * acc;
    }
}

// Two-pass separable form of the same adjoint for up-sampling factors (K >= 3 taps per axis): the K^3 gathers per coarse
// voxel of k_resize_bwd become K^2 + K.  Pass 1 reduces x and y at full z resolution (thread = (iy, z_fine) of one
// coarse x plane; every load is a coalesced row segment, the K^2 rows of neighbouring threads overlap in L1), pass 2
// reduces z.  Traffic: N_out read + N_out/4 written and read back + N_in written (x2 zoom), against N_out * K^3 / 8
// sector-granular gathers before.  Same weights, a different summation order than the one-pass kernel.
// block = 32 fine z x 8 coarse y; a thread walks RXS consecutive coarse x: the K rows of neighbouring iy and the K planes
// of consecutive ix overlap, and both overlaps are served by L1 (the first version, one (iy, z) row per block, re-read every
// fine element 6 times from L2 and ran at the one-pass kernel's speed)
constexpr int RXS = 4;
template <int K>
__global__ void __launch_bounds__(256)
k_resize_bwd_xy(const float *__restrict__ gout, float *__restrict__ mid, const int *__restrict__ xlo, const float *__restrict__ xw,
                int kx, const int *__restrict__ ylo, const float *__restrict__ yw, int ky, int Xi, int Yi, int Xo, int Yo, int Zo,
                int nxt) {
    const uint32_t z = blockIdx.x * 32u + (threadIdx.x & 31u), iy = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (z >= (uint32_t)Zo || iy >= (uint32_t)Yi) return;
    const uint32_t bc = blockIdx.z / (uint32_t)nxt, ix0 = (blockIdx.z - bc * (uint32_t)nxt) * RXS;
    const float *gb = gout + (size_t)bc * Xo * Yo * Zo + z;
    const int y0 = __ldg(ylo + iy);
    float wy[K];
    uint32_t oy[K];
#pragma unroll
    for (int b = 0; b < K; ++b) {
        wy[b] = b < ky ? __ldg(yw + iy * ky + b) : 0.f;
        oy[b] = (uint32_t)min(y0 + b, Yo - 1) * (uint32_t)Zo;
    }
    // (sharing each plane's y reduction between the outputs of the walk through a run-time plane loop was measured
    // slower -- 950 vs 733 us at B = 16: the loop serialises the loads that full unrolling keeps in flight together)
#pragma unroll
    for (int r = 0; r < RXS; ++r) {
        const uint32_t ix = ix0 + r;
        if (ix >= (uint32_t)Xi) break;
        const int x0 = __ldg(xlo + ix);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const float wx = a < kx ? __ldg(xw + ix * kx + a) : 0.f;
            const float *pl = gb + (size_t)min(x0 + a, Xo - 1) * Yo * Zo;
            float accy = 0.f;
#pragma unroll
            for (int b = 0; b < K; ++b) accy = fmaf(wy[b], __ldg(pl + oy[b]), accy);
            acc = fmaf(wx, accy, acc);
        }
        mid[(size_t)bc * Xi * Yi * Zo + ((size_t)ix * Yi + iy) * Zo + z] = acc;
    }
}

template <int K>
__global__ void __launch_bounds__(256)
k_resize_bwd_z(const float *__restrict__ mid, float *__restrict__ gin, const int *__restrict__ zlo, const float *__restrict__ zw, int kz,
               int Zi, int Zo, float s, FastDiv zdiv, size_t rows) {
    // thread = one coarse sample of one row; consecutive threads read overlapping K-wide windows of the same fine row
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= rows * (size_t)Zi) return;
    const uint32_t row = (uint32_t)(e / (uint32_t)Zi), iz = (uint32_t)(e - (size_t)row * Zi);
    (void)zdiv;
    const float *src = mid + (size_t)row * Zo;
    const int z0 = __ldg(zlo + iz);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const float w = c < kz ? __ldg(zw + iz * kz + c) : 0.f;
        acc = fmaf(w, __ldg(src + min(z0 + c, Zo - 1)), acc);
    }
    gin[e] = s * acc;
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_resize_fwd(const float *in, float *out, const float *cx, const float *cy, const float *cz,
                              int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo, float pre,
                              float post, int interp, unsigned flags, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xo >= 0 && Yo >= 0 && Zo >= 0, DFM_EINVAL,
                "dfm_resize_fwd: bad shape B=%d C=%d in=(%d,%d,%d) out=(%d,%d,%d)", B, C, Xi, Yi, Zi, Xo, Yo, Zo);
    DFM_REQUIRE(B <= 65535 && Xo <= 65535, DFM_EINVAL, "dfm_resize_fwd: B and Xo must be <= 65535");
    DFM_REQUIRE((uint64_t)Xo * Yo * Zo < (1ull << 30) && (uint64_t)Xi * Yi * Zi < (1ull << 30), DFM_EINVAL,
                "dfm_resize_fwd: volume too large (>= 2^30 voxels)");
    DFM_REQUIRE((uint64_t)Yo * Zo * (uint64_t)Zo < (1ull << 32), DFM_EINVAL, "dfm_resize_fwd: Yo*Zo*Zo must be < 2^32");
    DFM_REQUIRE(interp == DFM_LINEAR || interp == DFM_NEAREST, DFM_EINVAL, "dfm_resize_fwd: interp %d", interp);
    if (B == 0 || Xo == 0 || Yo == 0 || Zo == 0) return DFM_OK;
    DFM_REQUIRE(in && out && cx && cy && cz, DFM_EINVAL, "dfm_resize_fwd: null pointer");
    DFM_REQUIRE(in != out, DFM_EINVAL, "dfm_resize_fwd: out must not alias in");
    if (C == 1) flags &= ~(DFM_FIELD_IN_CL | DFM_FIELD_OUT_CL);
    cudaStream_t st = (cudaStream_t)stream;
    if (interp == DFM_LINEAR) {
        if (C == 3 && !(flags & (DFM_FIELD_IN_CL | DFM_FIELD_OUT_CL)) && Xo >= Xi && Yo >= Yi && Zo >= Zi) {
            int rc = launch_upsample3_march(in, out, cx, cy, cz, B, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
            rc = launch_resize3_smem(in, out, cx, cy, cz, B, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, st);
            if (rc != DFM_EUNSUPPORTED) return rc;
        }
        return launch_resize<DFM_LINEAR>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st);
    }
    return launch_resize<DFM_NEAREST>(in, out, cx, cy, cz, B, C, Xi, Yi, Zi, Xo, Yo, Zo, pre, post, flags, st);
}

extern "C" size_t dfm_resize_bwd_workspace_bytes(int B, int C, int Xi, int Yi, int Zo) {
    if (B <= 0 || C <= 0 || Xi <= 0 || Yi <= 0 || Zo <= 0) return 0;
    return (size_t)B * C * Xi * Yi * Zo * sizeof(float);
}

extern "C" int dfm_resize_bwd(const float *gout, float *gin, const int *xlo, const int *xcnt, const float *xw, int kx,
                              const int *ylo, const int *ycnt, const float *yw, int ky, const int *zlo, const int *zcnt,
                              const float *zw, int kz, int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo, int Zo,
                              float pre, float post, void *stream) {
    return dfm_resize_bwd_ws(gout, gin, nullptr, xlo, xcnt, xw, kx, ylo, ycnt, yw, ky, zlo, zcnt, zw, kz, B, C, Xi, Yi, Zi, Xo, Yo,
                             Zo, pre, post, stream);
}

extern "C" int dfm_resize_bwd_ws(const float *gout, float *gin, float *work, const int *xlo, const int *xcnt, const float *xw,
                                 int kx, const int *ylo, const int *ycnt, const float *yw, int ky, const int *zlo,
                                 const int *zcnt, const float *zw, int kz, int B, int C, int Xi, int Yi, int Zi, int Xo, int Yo,
                                 int Zo, float pre, float post, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 1 && Yi >= 1 && Zi >= 1 && Xo >= 0 && Yo >= 0 && Zo >= 0, DFM_EINVAL,
                "dfm_resize_bwd: bad shape");
    DFM_REQUIRE(kx >= 1 && ky >= 1 && kz >= 1, DFM_EINVAL, "dfm_resize_bwd: tap counts must be >= 1");
    DFM_REQUIRE((uint64_t)B * C <= 65535 && Xi <= 65535, DFM_EINVAL, "dfm_resize_bwd: B*C and Xi must be <= 65535");
    DFM_REQUIRE((uint64_t)Xo * Yo * Zo < (1ull << 31) && (uint64_t)Xi * Yi * Zi < (1ull << 31), DFM_EINVAL,
                "dfm_resize_bwd: volume too large");
    DFM_REQUIRE((uint64_t)Yi * Zi * (uint64_t)Zi < (1ull << 32), DFM_EINVAL, "dfm_resize_bwd: Yi*Zi*Zi must be < 2^32");
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(gout && gin && xlo && xcnt && xw && ylo && ycnt && yw && zlo && zcnt && zw, DFM_EINVAL,
                "dfm_resize_bwd: null pointer");
    (void)xcnt; (void)ycnt; (void)zcnt;   // table rows are zero-padded to k taps, so the counts are not needed on the device
    const uint32_t plane = (uint32_t)Yi * Zi;
    const int K = kx > ky ? (kx > kz ? kx : kz) : (ky > kz ? ky : kz);
    cudaStream_t st = (cudaStream_t)stream;
    const FastDiv fd = make_fastdiv(Zi);
    const float sc = pre * post;
    static const bool one_pass = getenv("DFM_RESIZE_BWD_ONEPASS") != nullptr;      // tuning / testing aid
    if (work && K >= 3 && !one_pass && (uint64_t)B * C * ((Xi + RXS - 1) / RXS) <= 65535) {      // separable two-pass adjoint
        const size_t rows = (size_t)B * C * Xi * Yi;
        const int nxt = (Xi + RXS - 1) / RXS;
        dim3 g1((Zo + 31) / 32, (Yi + 7) / 8, B * C * nxt);
        const unsigned g2 = (unsigned)((rows * Zi + 255) / 256);
#define DFM_SEP(KK)                                                                                                         \
    case KK:                                                                                                                \
        k_resize_bwd_xy<KK><<<g1, 256, 0, st>>>(gout, work, xlo, xw, kx, ylo, yw, ky, Xi, Yi, Xo, Yo, Zo, nxt);               \
        k_resize_bwd_z<KK><<<g2, 256, 0, st>>>(work, gin, zlo, zw, kz, Zi, Zo, sc, fd, rows);                               \
        break;
        switch (K) {
            DFM_SEP(3) DFM_SEP(4) DFM_SEP(5) DFM_SEP(6) DFM_SEP(7) DFM_SEP(8)
            default: return fail(DFM_EUNSUPPORTED, "dfm_resize_bwd: %d taps per input sample (max 8: zoom factors up to ~4)", K);
        }
#undef DFM_SEP
        return check_launch("dfm_resize_bwd(separable)");
    }
#define DFM_GO(KK, RX)                                                                                              \
    do {                                                                                                            \
        dim3 grid((plane + 255) / 256, (Xi + RX - 1) / RX, B * C), block(256);                                      \
        k_resize_bwd<KK, RX><<<grid, block, 0, st>>>(gout, gin, xlo, xw, kx, ylo, yw, ky, zlo, zw, kz, Xi, Yi, Zi, Xo, Yo, \
                                                     Zo, sc, fd, plane);                                            \
    } while (0)
    switch (K) {
        case 1: DFM_GO(1, 4); break;
        case 2: DFM_GO(2, 4); break;
        case 3: DFM_GO(3, 2); break;
        case 4: DFM_GO(4, 1); break;
        case 5: DFM_GO(5, 1); break;
        case 6: DFM_GO(6, 1); break;
        case 7: DFM_GO(7, 1); break;
        case 8: DFM_GO(8, 1); break;
        default: return fail(DFM_EUNSUPPORTED, "dfm_resize_bwd: %d taps per input sample (max 8: zoom factors up to ~4)", K);
    }
#undef DFM_GO
    return check_launch("dfm_resize_bwd");
}
