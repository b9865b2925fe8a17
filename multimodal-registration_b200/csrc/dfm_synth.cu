// Intensity model of neurite.models.labels_to_image [UR] (train_synthmorph.py:258-289),
// the voxel-level work that follows the label-map deformation (VecInt -> RescaleTransform -> nearest warp, the hot
// path's own kernels):
//   1. k_synth_intensity  image = mean[label] + std[label] * N(0, 1)      (per-label Gaussian intensities)
//   2. k_conv1d_axis      separable Gaussian blur, one pass per axis (zero padding, like Keras 'same')
//   3. k_scale_exp_clip   image * exp(bias field), clipped to [lo, hi]
//   4. k_norm_gamma       ((image - min) / (max - min)) ** gamma           (min / max from dfm_minmax, on the device)
//   5. k_onehot           label -> one-hot channels-last map through a lookup table (out_label_list)
// The random stream is counter-based (Philox-4x32-10, one counter per group of four voxels, Box-Muller): results are a
// pure function of (seed, voxel index), independent of the launch geometry.  It cannot match TensorFlow's stream, so
// parity for this row is distributional (tests/test_synth_gpu.py); everything downstream of the noise is deterministic.
#include <algorithm>

#include "dfm_common.cuh"

namespace dfm {

// Philox-4x32-10 (Salmon et al., SC'11): the round constants are the published ones
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

// one thread = four consecutive voxels = one Philox block = two Box-Muller pairs
__global__ void __launch_bounds__(256)
k_synth_intensity(const float *__restrict__ labels, const float *__restrict__ means, const float *__restrict__ stds, int nlabels,
                  uint64_t seed, float *__restrict__ out, size_t n) {
    const size_t g = blockIdx.x * 256ull + threadIdx.x, i0 = g * 4;
    if (i0 >= n) return;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float z[4];
    {
        const float ra = sqrtf(-2.f * __logf(u01(r.x))), rb = sqrtf(-2.f * __logf(u01(r.z)));
        float s, c;
        __sincosf(6.283185307179586f * u01(r.y), &s, &c);
        z[0] = ra * c; z[1] = ra * s;
        __sincosf(6.283185307179586f * u01(r.w), &s, &c);
        z[2] = rb * c; z[3] = rb * s;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t i = i0 + k;
        if (i >= n) break;
        const int lab = min(max((int)__ldg(labels + i), 0), nlabels - 1);
        out[i] = fmaf(__ldg(stds + lab), z[k], __ldg(means + lab));
    }
}

// out[.., p, ..] = sum_k taps[k] * in[.., p + k - K/2, ..] along `axis` (0: x, 1: y, 2: z), zero outside the volume
__global__ void __launch_bounds__(256)
k_conv1d_axis(const float *__restrict__ in, float *__restrict__ out, int X, int Y, int Z, int axis, const float *__restrict__ taps,
              int K, FastDiv zdiv, FastDiv ydiv) {
    const uint32_t N = (uint32_t)X * Y * Z;
    const uint32_t n = blockIdx.x * 256u + threadIdx.x;
    if (n >= N) return;
    const uint32_t q = fast_div(n, zdiv), z = n - q * zdiv.d, x = fast_div(q, ydiv), y = q - x * ydiv.d;
    const int pos = axis == 0 ? (int)x : axis == 1 ? (int)y : (int)z;
    const int len = axis == 0 ? X : axis == 1 ? Y : Z;
    const int stride = axis == 0 ? Y * Z : axis == 1 ? Z : 1;
    const float *src = in + (size_t)blockIdx.y * N + n;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const int d = k - K / 2, p = pos + d;
        if (p >= 0 && p < len) acc = fmaf(__ldg(taps + k), __ldg(src + (ptrdiff_t)d * stride), acc);
    }
    out[(size_t)blockIdx.y * N + n] = acc;
}

__global__ void __launch_bounds__(256)
k_scale_exp_clip(const float *__restrict__ img, const float *__restrict__ logbias, float *__restrict__ out, size_t n, float lo, float hi) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256ull) {
        const float b = logbias ? __expf(__ldg(logbias + i)) : 1.f;
        out[i] = fminf(fmaxf(__ldg(img + i) * b, lo), hi);
    }
}

// per item: ((v - min) / (max - min)) ** gamma[b]; minmax: B device {min, max} pairs (float64, from dfm_minmax)
__global__ void __launch_bounds__(256)
k_norm_gamma(const float *__restrict__ img, const double *__restrict__ minmax, const float *__restrict__ gamma, float *__restrict__ out,
             size_t n_per_item) {
    const int b = blockIdx.y;
    const float mn = (float)minmax[2 * b], mx = (float)minmax[2 * b + 1];
    const float inv = mx > mn ? 1.f / (mx - mn) : 0.f, gm = gamma ? __ldg(gamma + b) : 1.f;
    const float *src = img + (size_t)b * n_per_item;
    float *dst = out + (size_t)b * n_per_item;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n_per_item; i += (size_t)gridDim.x * 256ull) {
        const float v = (__ldg(src + i) - mn) * inv;
        dst[i] = v > 0.f ? __powf(v, gm) : 0.f;
    }
}

// out[i * C + c] = (lut[label[i]] == c); labels outside the table (or mapped to a negative entry) give an all-zero voxel
__global__ void __launch_bounds__(256)
k_onehot(const float *__restrict__ labels, const int *__restrict__ lut, int nlut, int C, float *__restrict__ out, size_t n, FastDiv cdiv) {
    const size_t e = blockIdx.x * 256ull + threadIdx.x;
    if (e >= n * (size_t)C) return;
    const uint32_t i = C == 1 ? (uint32_t)e : fast_div((uint32_t)e, cdiv), c = (uint32_t)e - i * (uint32_t)C;
    const int lab = (int)__ldg(labels + i);
    const int m = (lab >= 0 && lab < nlut) ? __ldg(lut + lab) : -1;
    __stcs(out + e, m == (int)c ? 1.f : 0.f);
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_synth_intensity(const float *labels, const float *means, const float *stds, int nlabels, uint64_t seed, float *out,
                                   size_t n, void *stream) {
    DFM_REQUIRE(labels && means && stds && out && nlabels >= 1, DFM_EINVAL, "dfm_synth_intensity: bad argument");
    if (n == 0) return DFM_OK;
    const size_t groups = (n + 3) / 4;
    k_synth_intensity<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, means, stds, nlabels, seed, out, n);
    return check_launch("k_synth_intensity");
}

extern "C" int dfm_conv1d_axis(const float *in, float *out, int B, int X, int Y, int Z, int axis, const float *taps, int K, void *stream) {
    DFM_REQUIRE(in && out && taps && in != out, DFM_EINVAL, "dfm_conv1d_axis: null or aliased pointer");
    DFM_REQUIRE(B >= 0 && B <= 65535 && X >= 1 && Y >= 1 && Z >= 1 && axis >= 0 && axis <= 2 && K >= 1 && (K & 1), DFM_EINVAL,
                "dfm_conv1d_axis: bad shape, axis or an even tap count");
    DFM_REQUIRE((uint64_t)X * Y * Z * (uint64_t)std::max(Y, Z) < (1ull << 32), DFM_EINVAL, "dfm_conv1d_axis: volume too large");
    if (B == 0) return DFM_OK;
    const uint32_t N = (uint32_t)X * Y * Z;
    k_conv1d_axis<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(in, out, X, Y, Z, axis, taps, K, make_fastdiv(Z), make_fastdiv(Y));
    return check_launch("k_conv1d_axis");
}

extern "C" int dfm_scale_exp_clip(const float *img, const float *logbias, float *out, size_t n, float lo, float hi, void *stream) {
    DFM_REQUIRE(img && out, DFM_EINVAL, "dfm_scale_exp_clip: null pointer");
    if (n == 0) return DFM_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_scale_exp_clip<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, logbias, out, n, lo, hi);
    return check_launch("k_scale_exp_clip");
}

extern "C" int dfm_norm_gamma(const float *img, const double *minmax, const float *gamma, float *out, int B, size_t n_per_item, void *stream) {
    DFM_REQUIRE(img && minmax && out && B >= 0 && B <= 65535, DFM_EINVAL, "dfm_norm_gamma: bad argument");
    if (B == 0 || n_per_item == 0) return DFM_OK;
    const unsigned blocks = (unsigned)std::min<size_t>((n_per_item + 255) / 256, 148 * 16);
    k_norm_gamma<<<dim3(blocks, B), 256, 0, (cudaStream_t)stream>>>(img, minmax, gamma, out, n_per_item);
    return check_launch("k_norm_gamma");
}

extern "C" int dfm_onehot(const float *labels, const int *lut, int nlut, int C, float *out, size_t n, void *stream) {
    DFM_REQUIRE(labels && lut && out && nlut >= 1 && C >= 1, DFM_EINVAL, "dfm_onehot: bad argument");
    DFM_REQUIRE((uint64_t)n * C * (uint64_t)C < (1ull << 32), DFM_EINVAL, "dfm_onehot: n * C * C must be < 2^32");
    if (n == 0) return DFM_OK;
    const size_t total = n * (size_t)C;
    k_onehot<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, lut, nlut, C, out, n, make_fastdiv(C));
    return check_launch("k_onehot");
}
