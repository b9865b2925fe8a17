// Channels-last multi-channel linear warp, forward and backward (the reference's own layout:
// SpatialTransformer('linear')([one_hot(26), flow]) at train_synthmorph.py:298).
//
// With the channels innermost, the C values of one corner are one contiguous run (104 bytes at
// C = 26), so a gather is a coalesced line read instead of 32 scattered words.  The kernels use
// the lanes of a warp for the CHANNELS:
//   phase A  lane = voxel: the warp's 32 consecutive voxels load their displacement (coalesced),
//            set up corner base / weights once and park them in shared memory;
//   phase B  lane = channel: for each of the 32 voxels the record is broadcast back (3-4 LDS.128)
//            and every lane gathers its channel of the 8 corners; small channel counts put
//            32 / Cpad voxels side by side in the warp, large ones loop over 32-channel chunks.
// Backward: d/d field sums over channels through shared memory (no atomics); d/d img scatters
// with coalesced float atomics (reference semantics: gather back-propagates as scatter-add).
// Arithmetic per value is the shared tri_weights / tri_accumulate, so the forward pass is
// bit-identical to the other linear kernels in both builds.
#include "dfm_common.cuh"

namespace dfm {

struct __align__(16) FwdRec {
    float w[8];
    uint32_t base;      // element index of corner (0,0,0), channel 0
    uint32_t dead;      // fill_value applies
    uint32_t pad[2];
};

// opaque move: keeps a warp-uniform value in a vector register (the compiler would otherwise hold it
// in a uniform register, which the 64-bit address arithmetic cannot take as the 32-bit multiplicand)
__device__ __forceinline__ uint32_t vreg(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

__device__ __forceinline__ void voxel_xyz(uint32_t n, FastDiv zdiv, FastDiv ydiv, uint32_t &x, uint32_t &y, uint32_t &z) {
    const uint32_t q = fast_div(n, zdiv);
    z = n - q * zdiv.d;
    x = fast_div(q, ydiv);
    y = q - x * ydiv.d;
}

// V = 1 or 2 consecutive channels per lane (V = 2 needs an even C: every run then starts 8-byte aligned)
template <int V> __device__ __forceinline__ void ldv(const float *p, float (&a)[V]);
template <> __device__ __forceinline__ void ldv<1>(const float *p, float (&a)[1]) { a[0] = __ldg(p); }
template <> __device__ __forceinline__ void ldv<2>(const float *p, float (&a)[2]) {
    const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    a[0] = t.x; a[1] = t.y;
}
// streaming variants (read once / written once: keep them out of the L1 the corner gathers live in)
template <int V> __device__ __forceinline__ void ldv_stream(const float *p, float (&a)[V]);
template <> __device__ __forceinline__ void ldv_stream<1>(const float *p, float (&a)[1]) { a[0] = __ldcs(p); }
template <> __device__ __forceinline__ void ldv_stream<2>(const float *p, float (&a)[2]) {
    const float2 t = __ldcs(reinterpret_cast<const float2 *>(p));
    a[0] = t.x; a[1] = t.y;
}
template <int V> __device__ __forceinline__ void stv(float *p, const float (&a)[V]);
template <> __device__ __forceinline__ void stv<1>(float *p, const float (&a)[1]) { *p = a[0]; }
template <> __device__ __forceinline__ void stv<2>(float *p, const float (&a)[2]) { *reinterpret_cast<float2 *>(p) = make_float2(a[0], a[1]); }

// phase A shared by forward and backward: sample location of voxel n (lane = voxel)
__device__ __forceinline__ void load_loc(const float *__restrict__ fb, uint32_t n, uint32_t N, bool field_cl, bool absolute,
                                         FastDiv zdiv, FastDiv ydiv, float &lx, float &ly, float &lz) {
    if (field_cl) { lx = __ldg(fb + (size_t)n * 3); ly = __ldg(fb + (size_t)n * 3 + 1); lz = __ldg(fb + (size_t)n * 3 + 2); }
    else { lx = __ldg(fb + n); ly = __ldg(fb + N + n); lz = __ldg(fb + 2 * (size_t)N + n); }
    if (!absolute) {
        uint32_t x, y, z;
        voxel_xyz(n, zdiv, ydiv, x, y, z);
        lx = __fadd_rn((float)x, lx); ly = __fadd_rn((float)y, ly); lz = __fadd_rn((float)z, lz);
    }
}

// CSHIFT: log2 of the lanes per voxel (Cpad >= C, or 32 with MULTI = loop over 32-channel chunks)
template <int CSHIFT, int V, bool MULTI, bool HF>
__global__ void __launch_bounds__(256)
k_warp_cl(const float *__restrict__ img, const float *__restrict__ field, float *__restrict__ out, int C, int Xi,
          int Yi, int Zi, uint32_t N, float fill, int field_cl, int absolute, FastDiv zdiv, FastDiv ydiv) {
    __shared__ FwdRec s_rec[8][32];
    constexpr int CPAD = 1 << CSHIFT, VPW = 32 >> CSHIFT;           // voxels side by side in the warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n0 = blockIdx.x * 256u + warp * 32u;          // first voxel of this warp
    if (n0 >= N) return;
    const uint32_t Ni = (uint32_t)Xi * Yi * Zi;
    const float *ib = img + (size_t)blockIdx.y * C * Ni;
    {   // ---- phase A: lane = voxel ---------------------------------------------------------
        const uint32_t n = min(n0 + lane, N - 1);
        float lx, ly, lz;
        load_loc(field + (size_t)blockIdx.y * 3 * N, n, N, field_cl, absolute, zdiv, ydiv, lx, ly, lz);
        const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
        const AxisF ax = axis_fast(lx, (float)mxi, mxi), ay = axis_fast(ly, (float)myi, myi), az = axis_fast(lz, (float)mzi, mzi);
        FwdRec r;
        tri_weights(ax, ay, az, r.w);
        r.base = (((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1)) * (uint32_t)C;
        r.dead = HF && (lx < 0.f || lx > (float)mxi || ly < 0.f || ly > (float)myi || lz < 0.f || lz > (float)mzi);
        r.pad[0] = r.pad[1] = 0;
        s_rec[warp][lane] = r;
    }
    __syncwarp();
    // ---- phase B: lane = V channels ------------------------------------------------------------
    const int sub = lane >> CSHIFT, c0 = (lane & (CPAD - 1)) * V;
    const uint32_t oC = (uint32_t)C, oZ = (uint32_t)Zi * C, oY = (uint32_t)Yi * Zi * C;
    const uint32_t o3 = oZ + oC, o5 = oY + oC, o6 = oY + oZ, o7 = oY + oZ + oC;
    const bool lane_on = MULTI || c0 < C;
    const uint32_t nv = min(32u, N - n0);                             // voxels of this warp (warp-uniform)
    float *po = out + (size_t)blockIdx.y * C * N + (size_t)n0 * C + (uint32_t)sub * oC + c0;
    const float *pi = ib + c0;
#pragma unroll 4
    for (int j0 = 0; j0 < 32; j0 += VPW, po += VPW * oC) {
        const int j = j0 + sub;
        if ((uint32_t)j >= nv || !lane_on) continue;
        const FwdRec r = s_rec[warp][j];
        const float *p = pi + r.base;
        for (int cc = 0; cc < (MULTI ? C - c0 : 1); cc += 32 * V) {
            const float *q = p + cc;
            float v[8][V];
            ldv<V>(q, v[0]); ldv<V>(q + oC, v[1]); ldv<V>(q + oZ, v[2]); ldv<V>(q + o3, v[3]);
            ldv<V>(q + oY, v[4]); ldv<V>(q + o5, v[5]); ldv<V>(q + o6, v[6]); ldv<V>(q + o7, v[7]);
            float a[V];
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const float t[8] = {v[0][u], v[1][u], v[2][u], v[3][u], v[4][u], v[5][u], v[6][u], v[7][u]};
                a[u] = tri_accumulate(r.w, t);
                if (HF && r.dead) a[u] = fill;
            }
            stv<V>(po + cc, a);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct __align__(16) BwdRec {
    float dx[4];        // m_x * wy[b] * wz[c]   (index b*2+c): weight of (v[1,b,c] - v[0,b,c])
    float dy[4];        // m_y * wx[a] * wz[c]   (index a*2+c)
    float dz[4];        // m_z * wx[a] * wy[b]   (index a*2+b)
    uint32_t base;
    uint32_t pad[3];
};

// GDICE: `gout` is the map y_true and the upstream gradient is formed on the fly (fusing the Dice SUMS into the
// forward kernel the same way was measured and lost: a DRAM-latency read inside the gather loop costs 1.11 ms
// against 0.64 + 0.35 ms for the warp and a separate streaming pass),
// g[n, c] = coef[b][c][0] * y_true[n, c] + coef[b][c][1] (the Dice gradient, dfm_dice_bwd), never materialised.
// 4 CTAs/SM (64 registers): the deeper load batching beats the extra warps of a tighter register budget
// (1.07 ms vs 1.16 ms at 6 CTAs/SM; staging the once-read map in shared memory first: 1.27 ms).
template <int CSHIFT, int V, bool MULTI, bool NEED_IMG, bool NEED_FIELD, bool GDICE = false>
__global__ void __launch_bounds__(256, (NEED_IMG || MULTI) ? 1 : 4)
k_warp_cl_bwd(const float *__restrict__ gout, const float *__restrict__ img, const float *__restrict__ field,
              float *__restrict__ gimg, float *__restrict__ gfield, int C, int Xi, int Yi, int Zi, uint32_t N,
              int has_fill, int field_cl, int gfield_cl, FastDiv zdiv, FastDiv ydiv,
              const float *__restrict__ coef = nullptr) {
    static_assert(!(GDICE && MULTI), "fused Dice gradient needs fixed channels per lane");
    constexpr int CPAD = 1 << CSHIFT, VPW = 32 >> CSHIFT;
    constexpr int CV = VPW > 8 ? VPW : (CSHIFT == 5 ? 4 : 8), RL = CPAD + 1;          // reduction chunk (voxels), row pitch
    __shared__ BwdRec s_rec[8][32];
    __shared__ __align__(16) float s_w[NEED_IMG ? 8 : 1][32][8];
    __shared__ float s_red[NEED_FIELD ? 8 : 1][3 * CV * RL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n0 = blockIdx.x * 256u + warp * 32u;
    if (n0 >= N) return;
    const uint32_t Ni = (uint32_t)Xi * Yi * Zi;
    {
        const uint32_t n = min(n0 + lane, N - 1);
        float lx, ly, lz;
        load_loc(field + (size_t)blockIdx.y * 3 * N, n, N, field_cl, false, zdiv, ydiv, lx, ly, lz);
        const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
        const float mxf = (float)mxi, myf = (float)myi, mzf = (float)mzi;
        const AxisF ax = axis_fast(lx, mxf, mxi), ay = axis_fast(ly, myf, myi), az = axis_fast(lz, mzf, mzi);
        const bool dead = has_fill && (lx < 0.f || lx > mxf || ly < 0.f || ly > myf || lz < 0.f || lz > mzf);
        // clip passes the gradient on [0, max]; at loc == max both reference corners are the edge voxel
        // (difference 0), which the i1 - 1 addressing reproduces with a strict upper bound
        const float gx = (!dead && lx >= 0.f && lx < mxf) ? 1.f : 0.f;
        const float gy = (!dead && ly >= 0.f && ly < myf) ? 1.f : 0.f;
        const float gz = (!dead && lz >= 0.f && lz < mzf) ? 1.f : 0.f;
        const float wx[2] = {ax.w0, ax.w1}, wy[2] = {ay.w0, ay.w1}, wz[2] = {az.w0, az.w1};
        BwdRec r;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                r.dx[p * 2 + q] = gx * wy[p] * wz[q];
                r.dy[p * 2 + q] = gy * wx[p] * wz[q];
                r.dz[p * 2 + q] = gz * wx[p] * wy[q];
            }
        r.base = (((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1)) * (uint32_t)C;
        r.pad[0] = r.pad[1] = r.pad[2] = 0;
        s_rec[warp][lane] = r;
        if (NEED_IMG) {
            float w[8];
            tri_weights(ax, ay, az, w);
#pragma unroll
            for (int k = 0; k < 8; ++k) s_w[warp][lane][k] = dead ? 0.f : w[k];
        }
    }
    __syncwarp();
    const int sub = lane >> CSHIFT, c0 = (lane & (CPAD - 1)) * V;
    const uint32_t oC = (uint32_t)C, oZ = (uint32_t)Zi * C, oY = (uint32_t)Yi * Zi * C;
    const uint32_t o3 = oZ + oC, o5 = oY + oC, o6 = oY + oZ, o7 = oY + oZ + oC;
    const bool lane_on = MULTI || c0 < C;
    const int nl = min(CPAD, (C + V - 1) / V);                        // lanes of a voxel that carry channels
    const uint32_t nv = min(32u, N - n0);
    const float *pg = gout + (size_t)blockIdx.y * C * N + (size_t)n0 * C + (uint32_t)sub * oC + c0;
    const float *pi = img + (size_t)blockIdx.y * C * Ni + c0;
    float *pq = NEED_IMG ? gimg + (size_t)blockIdx.y * C * Ni + c0 : nullptr;
    float ca[V], cb[V];                                      // GDICE: coefficients of this lane's channels
#pragma unroll
    for (int u = 0; u < V; ++u) {
        ca[u] = (GDICE && lane_on) ? __ldg(coef + ((size_t)blockIdx.y * C + c0 + u) * 2) : 0.f;
        cb[u] = (GDICE && lane_on) ? __ldg(coef + ((size_t)blockIdx.y * C + c0 + u) * 2 + 1) : 0.f;
    }
    float mine[3] = {0.f, 0.f, 0.f};                       // gradient of voxel n0 + lane
#pragma unroll 4
    for (int j0 = 0; j0 < 32; j0 += VPW, pg += VPW * oC) {
        const int j = j0 + sub;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if ((uint32_t)j < nv && lane_on) {
            const BwdRec r = s_rec[warp][j];
            for (int cc = 0; cc < (MULTI ? C - c0 : 1); cc += 32 * V) {
                float g[V];
                ldv_stream<V>(pg + cc, g);
                if (GDICE) {
#pragma unroll
                    for (int u = 0; u < V; ++u) g[u] = fmaf(ca[u], g[u], cb[u]);
                }
                if (NEED_FIELD) {
                    const float *p = pi + r.base + cc;
                    float v[8][V];
                    ldv<V>(p, v[0]); ldv<V>(p + oC, v[1]); ldv<V>(p + oZ, v[2]); ldv<V>(p + o3, v[3]);
                    ldv<V>(p + oY, v[4]); ldv<V>(p + o5, v[5]); ldv<V>(p + o6, v[6]); ldv<V>(p + o7, v[7]);
#pragma unroll
                    for (int u = 0; u < V; ++u) {
                        float dx = r.dx[0] * (v[4][u] - v[0][u]);
                        dx = fmaf(r.dx[1], v[5][u] - v[1][u], dx); dx = fmaf(r.dx[2], v[6][u] - v[2][u], dx); dx = fmaf(r.dx[3], v[7][u] - v[3][u], dx);
                        float dy = r.dy[0] * (v[2][u] - v[0][u]);
                        dy = fmaf(r.dy[1], v[3][u] - v[1][u], dy); dy = fmaf(r.dy[2], v[6][u] - v[4][u], dy); dy = fmaf(r.dy[3], v[7][u] - v[5][u], dy);
                        float dz = r.dz[0] * (v[1][u] - v[0][u]);
                        dz = fmaf(r.dz[1], v[3][u] - v[2][u], dz); dz = fmaf(r.dz[2], v[5][u] - v[4][u], dz); dz = fmaf(r.dz[3], v[7][u] - v[6][u], dz);
                        sx = fmaf(g[u], dx, sx); sy = fmaf(g[u], dy, sy); sz = fmaf(g[u], dz, sz);
                    }
                }
                if (NEED_IMG) {
                    float *q = pq + r.base + cc;
                    const float4 wa = *reinterpret_cast<const float4 *>(&s_w[warp][j][0]);
                    const float4 wb = *reinterpret_cast<const float4 *>(&s_w[warp][j][4]);
                    const uint32_t off[8] = {0u, oC, oZ, o3, oY, o5, o6, o7};
                    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k)
#pragma unroll
                        for (int u = 0; u < V; ++u) {
                            const float a = w[k] * g[u];
                            if (a != 0.f) atomicAdd(q + off[k] + u, a);
                        }
                }
            }
        }
        if (NEED_FIELD) {
            // channel sum through shared memory: each lane parks its partial sums in the row of its
            // voxel; after a chunk of CV voxels the lanes owning those voxels add up their rows
            // (odd row pitch: conflict-free both ways) -- about 4 instructions per voxel, no shuffles
            const int jj = j & (CV - 1);
            if ((uint32_t)j < nv && lane_on) {
                float *rr = &s_red[warp][jj * RL + (lane & (CPAD - 1))];
                rr[0] = sx; rr[CV * RL] = sy; rr[2 * CV * RL] = sz;
            }
            if (((j0 + VPW) & (CV - 1)) == 0) {                        // chunk complete (compile-time after unrolling)
                __syncwarp();
                const int base = j0 + VPW - CV;
                if (lane >= base && lane < base + CV) {
                    const float *rr = &s_red[warp][(lane - base) * RL];
                    float ax = 0.f, ay = 0.f, az = 0.f;
                    for (int i = 0; i < nl; ++i) { ax += rr[i]; ay += rr[CV * RL + i]; az += rr[2 * CV * RL + i]; }
                    mine[0] = ax; mine[1] = ay; mine[2] = az;
                }
                __syncwarp();
            }
        }
    }
    if (NEED_FIELD) {
        const uint32_t n = n0 + lane;
        if (n < N) {
            float *gf = gfield + (size_t)blockIdx.y * 3 * N;
            if (gfield_cl) { gf[(size_t)n * 3] = mine[0]; gf[(size_t)n * 3 + 1] = mine[1]; gf[(size_t)n * 3 + 2] = mine[2]; }
            else { gf[n] = mine[0]; gf[N + n] = mine[1]; gf[2 * (size_t)N + n] = mine[2]; }
        }
    }
}

// ------------------------------- host side -----------------------------------------------
static bool cl_ok(int C, int Xi, int Yi, int Zi) {
    static const bool off = getenv("DFM_NO_WARP_CL") != nullptr;          // tuning aid
    return !off && C > 1 && Xi >= 2 && Yi >= 2 && Zi >= 2 && (uint64_t)Xi * Yi * Zi * (uint64_t)C < (1ull << 31);
}
static int cpad_shift(int C) {
    int s = 0;
    while ((1 << s) < C && s < 5) ++s;
    return s;
}
static bool aligned8(const void *p) { return ((uintptr_t)p & 7u) == 0; }

int launch_warp_cl_fwd(const float *img, const float *field, float *out, int B, int C, int Xi, int Yi, int Zi, int X,
                       int Y, int Z, int has_fill, float fill, unsigned flags, cudaStream_t st) {
    if (!cl_ok(C, Xi, Yi, Zi) || (uint64_t)X * Y * Z * (uint64_t)C >= (1ull << 31)) return DFM_EUNSUPPORTED;
    const uint32_t N = (uint32_t)X * Y * Z;
    dim3 grid((N + 255) / 256, B), block(256);
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    const int fcl = (flags & DFM_FIELD_IN_CL) ? 1 : 0, ab = (flags & DFM_LOC_ABSOLUTE) ? 1 : 0;
    static const bool no_v2 = getenv("DFM_CL_NO_V2") != nullptr;      // tuning aid
    const bool v2 = !no_v2 && C >= 4 && C % 2 == 0 && aligned8(img) && aligned8(out);
    const int lanes = v2 ? C / 2 : C;                                 // lanes a voxel needs
#define DFM_GO(S, V, M, H) k_warp_cl<S, V, M, H><<<grid, block, 0, st>>>(img, field, out, C, Xi, Yi, Zi, N, fill, fcl, ab, zd, yd)
#define DFM_GO2(S, V, M) do { if (has_fill) DFM_GO(S, V, M, true); else DFM_GO(S, V, M, false); } while (0)
#define DFM_GO3(V)                                      \
    if (lanes > 32) DFM_GO2(5, V, true);                \
    else switch (cpad_shift(lanes)) {                   \
        case 0: case 1: DFM_GO2(1, V, false); break;    \
        case 2: DFM_GO2(2, V, false); break;            \
        case 3: DFM_GO2(3, V, false); break;            \
        case 4: DFM_GO2(4, V, false); break;            \
        default: DFM_GO2(5, V, false); break;           \
    }
    if (v2) { DFM_GO3(2) } else { DFM_GO3(1) }
#undef DFM_GO3
#undef DFM_GO2
#undef DFM_GO
    return check_launch("k_warp_cl");
}

int launch_warp_cl_bwd(const float *gout, const float *img, const float *field, float *gimg, float *gfield, int B,
                       int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags,
                       cudaStream_t st) {
    if (!cl_ok(C, Xi, Yi, Zi) || (uint64_t)X * Y * Z * (uint64_t)C >= (1ull << 31)) return DFM_EUNSUPPORTED;
    const uint32_t N = (uint32_t)X * Y * Z;
    dim3 grid((N + 255) / 256, B), block(256);
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    const int fcl = (flags & DFM_FIELD_IN_CL) ? 1 : 0, gcl = (flags & DFM_FIELD_OUT_CL) ? 1 : 0;
    static const bool no_v2 = getenv("DFM_CL_NO_V2") != nullptr;      // tuning aid
    const bool v2 = !no_v2 && C >= 4 && C % 2 == 0 && aligned8(img) && aligned8(gout) && aligned8(gimg);
    const int lanes = v2 ? C / 2 : C;
#define DFM_GO(S, V, M, I, G) k_warp_cl_bwd<S, V, M, I, G><<<grid, block, 0, st>>>(gout, img, field, gimg, gfield, C, Xi, Yi, Zi, N, has_fill, fcl, gcl, zd, yd)
#define DFM_GO2(S, V, M) do { if (gimg && gfield) DFM_GO(S, V, M, true, true); else if (gimg) DFM_GO(S, V, M, true, false); else DFM_GO(S, V, M, false, true); } while (0)
#define DFM_GO3(V)                                      \
    if (lanes > 32) DFM_GO2(5, V, true);                \
    else switch (cpad_shift(lanes)) {                   \
        case 0: case 1: DFM_GO2(1, V, false); break;    \
        case 2: DFM_GO2(2, V, false); break;            \
        case 3: DFM_GO2(3, V, false); break;            \
        case 4: DFM_GO2(4, V, false); break;            \
        default: DFM_GO2(5, V, false); break;           \
    }
    if (v2) { DFM_GO3(2) } else { DFM_GO3(1) }
#undef DFM_GO3
#undef DFM_GO2
#undef DFM_GO
    return check_launch("k_warp_cl_bwd");
}

// ------------------------------- fused Dice ------------------------------------------------
static bool dice_shape_ok(int C, bool v2) { return v2 ? (C % 2 == 0 && C <= 64) : C <= 32; }

int launch_warp_cl_dice_bwd(const float *y_true, const float *coef, const float *img, const float *field, float *gfield,
                            int B, int C, int Xi, int Yi, int Zi, int X, int Y, int Z, int has_fill, unsigned flags,
                            cudaStream_t st) {
    if (!cl_ok(C, Xi, Yi, Zi) || (uint64_t)X * Y * Z * (uint64_t)C >= (1ull << 31)) return DFM_EUNSUPPORTED;
    const bool v2 = C >= 4 && C % 2 == 0 && aligned8(img) && aligned8(y_true);
    if (!dice_shape_ok(C, v2)) return DFM_EUNSUPPORTED;
    const uint32_t N = (uint32_t)X * Y * Z;
    dim3 grid((N + 255) / 256, B), block(256);
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    const int fcl = (flags & DFM_FIELD_IN_CL) ? 1 : 0, gcl = (flags & DFM_FIELD_OUT_CL) ? 1 : 0;
    const int lanes = v2 ? C / 2 : C;
#define DFM_GO(S, V) k_warp_cl_bwd<S, V, false, false, true, true><<<grid, block, 0, st>>>(y_true, img, field, nullptr, gfield, C, Xi, Yi, Zi, N, has_fill, fcl, gcl, zd, yd, coef)
#define DFM_GO3(V)                                   \
    switch (cpad_shift(lanes)) {                     \
        case 0: case 1: DFM_GO(1, V); break;         \
        case 2: DFM_GO(2, V); break;                 \
        case 3: DFM_GO(3, V); break;                 \
        case 4: DFM_GO(4, V); break;                 \
        default: DFM_GO(5, V); break;                \
    }
    if (v2) { DFM_GO3(2) } else { DFM_GO3(1) }
#undef DFM_GO3
#undef DFM_GO
    return check_launch("k_warp_cl_bwd(dice)");
}

}  // namespace dfm
