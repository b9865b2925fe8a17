// Channel-wise linear warp: every channel of a channels-last volume moves by its OWN 3-vector field,
//   out[b, p, c] = interp(vol[b, ..., c], p + shift[b, p, c, :])
// -- vxm.utils.transform(im, warp) of generate_label_maps (train_synthmorph.py:61-67): `im` is a
// [X, Y, Z, 26] noise volume, `warp` the [X, Y, Z, 26, 3] Perlin field (1.53 GB at 160 x 160 x 192), and the
// result goes straight into tf.argmax(im, axis=-1) (:68).
//
// The reference evaluates it as a 4-D interpn with the channel index as an extra integral coordinate
// (SURVEY.md Appendix A.3): the channel weights are exactly 0 / 1, so each value is the 8-corner spatial
// sum in corner order (the interleaved zero terms add exactly 0).  Both tensors are addressed in place in the
// reference's own layout -- no transposition passes: the flat element index e = (b, p, c) is the lane index,
// so the 12-byte shift records and the outputs of a warp are contiguous, and the 8 corner reads of
// neighbouring channels coincide in voxel wherever neighbouring channels move alike (draw_perlin samples the
// label axis at ceil(26 / scale) points: they nearly do).
// ARGMAX: the reduction over channels (first maximum, like tf.argmax) is taken inside the warp and a uint8
// label map is written instead of the C-channel result (-4C bytes per voxel).
#include <algorithm>

#include "dfm_common.cuh"

namespace dfm {

template <bool HF>
__device__ __forceinline__ float cw_sample(const float *__restrict__ vb, const float *__restrict__ sb, uint32_t e, uint32_t c,
                                           uint32_t x, uint32_t y, uint32_t z, int C, int Xi, int Yi, int Zi, float fill) {
    const float *s = sb + (size_t)e * 3;
    const float lx = __fadd_rn((float)x, __ldcs(s)), ly = __fadd_rn((float)y, __ldcs(s + 1)), lz = __fadd_rn((float)z, __ldcs(s + 2));
    const int mxi = Xi - 1, myi = Yi - 1, mzi = Zi - 1;
    const AxisF ax = axis_fast(lx, (float)mxi, mxi), ay = axis_fast(ly, (float)myi, myi), az = axis_fast(lz, (float)mzi, mzi);
    float w[8], v[8];
    tri_weights(ax, ay, az, w);
    const uint32_t lo = (((uint32_t)(ax.i1 - 1) * Yi + (uint32_t)(ay.i1 - 1)) * Zi + (uint32_t)(az.i1 - 1));
    gather8(vb + (size_t)lo * C + c, (uint32_t)Zi * C, (uint32_t)Yi * Zi * C, (uint32_t)C, v);
    float r = tri_accumulate(w, v);
    if (HF && (lx < 0.f || lx > (float)mxi || ly < 0.f || ly > (float)myi || lz < 0.f || lz > (float)mzi)) r = fill;
    return r;
}

// one thread per output element, elements in memory order (channel fastest)
template <bool HF>
__global__ void __launch_bounds__(256)
k_warp_cw(const float *__restrict__ vol, const float *__restrict__ shift, float *__restrict__ out, int C, int Xi, int Yi,
          int Zi, uint32_t N, float fill, FastDiv cdiv, FastDiv zdiv, FastDiv ydiv) {
    const uint32_t NC = N * (uint32_t)C;
    const uint32_t e = blockIdx.x * 256u + threadIdx.x;
    if (e >= NC) return;
    const uint32_t n = fast_div(e, cdiv), c = e - n * cdiv.d;
    const uint32_t q = fast_div(n, zdiv), z = n - q * zdiv.d, x = fast_div(q, ydiv), y = q - x * ydiv.d;
    const size_t Ni = (size_t)Xi * Yi * Zi;
    const float r = cw_sample<HF>(vol + (size_t)blockIdx.y * Ni * C, shift + (size_t)blockIdx.y * NC * 3, e, c, x, y, z, C, Xi, Yi, Zi, fill);
    __stcs(out + (size_t)blockIdx.y * NC + e, r);
}

// ARGMAX: a block owns VPB = 256 / C whole voxels (thread = one (voxel, channel) element in memory order, so the
// loads are as contiguous as in k_warp_cw); the values meet in shared memory and one thread per voxel scans its C
// channels in ascending order with `>` -- exactly tf.argmax's first maximum.
template <bool HF>
__global__ void __launch_bounds__(256)
k_warp_cw_argmax(const float *__restrict__ vol, const float *__restrict__ shift, uint8_t *__restrict__ lab, int C, int Xi,
                 int Yi, int Zi, uint32_t N, float fill, FastDiv cdiv, FastDiv zdiv, FastDiv ydiv, int vpb) {
    __shared__ float s_val[256];
    const uint32_t v = fast_div(threadIdx.x, cdiv), c = threadIdx.x - v * cdiv.d;
    const uint32_t n = blockIdx.x * (uint32_t)vpb + v;
    if (v < (uint32_t)vpb && n < N) {
        const uint32_t q = fast_div(n, zdiv), z = n - q * zdiv.d, x = fast_div(q, ydiv), y = q - x * ydiv.d;
        const size_t Ni = (size_t)Xi * Yi * Zi;
        s_val[threadIdx.x] = cw_sample<HF>(vol + (size_t)blockIdx.y * Ni * C, shift + (size_t)blockIdx.y * (size_t)N * C * 3,
                                           n * (uint32_t)C + c, c, x, y, z, C, Xi, Yi, Zi, fill);
    }
    __syncthreads();
    if (threadIdx.x < (uint32_t)vpb) {
        const uint32_t m = blockIdx.x * (uint32_t)vpb + threadIdx.x;
        if (m < N) {
            const float *p = s_val + threadIdx.x * C;
            float bv = p[0];
            int bi = 0;
            for (int k = 1; k < C; ++k)
                if (p[k] > bv) { bv = p[k]; bi = k; }
            lab[(size_t)blockIdx.y * N + m] = (uint8_t)bi;
        }
    }
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_warp_channelwise_fwd(const float *vol, const float *shift, void *out, int B, int C, int Xi, int Yi, int Zi,
                                        int X, int Y, int Z, int has_fill, float fill, int argmax, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && Xi >= 2 && Yi >= 2 && Zi >= 2 && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL,
                "dfm_warp_channelwise_fwd: bad shape (every volume axis must be >= 2)");
    DFM_REQUIRE(B <= 65535, DFM_EINVAL, "dfm_warp_channelwise_fwd: B must be <= 65535");
    DFM_REQUIRE((uint64_t)X * Y * Z * C < (1ull << 32) && (uint64_t)Xi * Yi * Zi * C < (1ull << 32), DFM_EINVAL,
                "dfm_warp_channelwise_fwd: X*Y*Z*C must be < 2^32");
    DFM_REQUIRE((uint64_t)X * Y * Z * (uint64_t)std::max(std::max(C * C, Z), Y) < (1ull << 32), DFM_EUNSUPPORTED,
                "dfm_warp_channelwise_fwd: volume too large for the 32-bit index arithmetic (X*Y*Z*C*C must be < 2^32)");
    DFM_REQUIRE(!argmax || C <= 256, DFM_EUNSUPPORTED, "dfm_warp_channelwise_fwd: argmax needs C <= 256 (uint8 labels; got %d)", C);
    if (B == 0) return DFM_OK;
    DFM_REQUIRE(vol && shift && out, DFM_EINVAL, "dfm_warp_channelwise_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t N = (uint32_t)X * Y * Z;
    const FastDiv zd = make_fastdiv(Z), yd = make_fastdiv(Y);
    if (!argmax) {
        const uint64_t NC = (uint64_t)N * C;
        dim3 grid((unsigned)((NC + 255) / 256), B);
        if (has_fill) k_warp_cw<true><<<grid, 256, 0, st>>>(vol, shift, (float *)out, C, Xi, Yi, Zi, N, fill, make_fastdiv(C), zd, yd);
        else k_warp_cw<false><<<grid, 256, 0, st>>>(vol, shift, (float *)out, C, Xi, Yi, Zi, N, fill, make_fastdiv(C), zd, yd);
        return check_launch("k_warp_cw");
    }
    const int vpb = 256 / C;
    dim3 grid((N + vpb - 1) / vpb, B);
    if (has_fill) k_warp_cw_argmax<true><<<grid, 256, 0, st>>>(vol, shift, (uint8_t *)out, C, Xi, Yi, Zi, N, fill, make_fastdiv(C), zd, yd, vpb);
    else k_warp_cw_argmax<false><<<grid, 256, 0, st>>>(vol, shift, (uint8_t *)out, C, Xi, Yi, Zi, N, fill, make_fastdiv(C), zd, yd, vpb);
    return check_launch("k_warp_cw_argmax");
}
