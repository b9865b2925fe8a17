// Voxel-level evaluation metrics of the reference's eval scripts:
//   * joint histogram for the normalised mutual information of eval_reg_with_mi.py:38-74
//     (np.histogramdd([a, b], bins) semantics: per-image edges = np.linspace(min, max, bins + 1) in float64,
//     bin = searchsorted(edges, v, 'right') - 1, the right-most edge belongs to the last bin);
//   * min / max of an image (the edges' end points);
//   * per-axis plane sums for detect_zero_padding (eval_reg_with_mi.py:16-36);
//   * the masked sums behind the overlap metrics of eval_reg_on_sc_seg.py:80-124 (TP / FP / TN / FN).
// All streaming, HBM-bound, one pass over the volumes; counts are integers (order-independent); the overlap sums
// are reduced in float64 in a fixed order (deterministic), the plane sums (only ever compared with 0) by float64 atomics.
#include <float.h>
#include <algorithm>

#include "dfm_common.cuh"

namespace dfm {

template <typename T>
__device__ __forceinline__ double ldd(const void *p, size_t i) { return (double)__ldg(reinterpret_cast<const T *>(p) + i); }

// ---- min / max ------------------------------------------------------------------------------
// out[2 * blockIdx.x] = {min, max} of the block's grid-stride share; a second launch with one block folds them
template <typename T>
__global__ void __launch_bounds__(256)
k_minmax(const void *__restrict__ a, size_t n, double *__restrict__ out) {
    __shared__ double s_mn[8], s_mx[8];
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256ull) {
        const double v = ldd<T>(a, i);
        mn = fmin(mn, v); mx = fmax(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fmin(mn, s_mn[w]); mx = fmax(mx, s_mx[w]); }
        out[2 * blockIdx.x] = mn; out[2 * blockIdx.x + 1] = mx;
    }
}

__global__ void k_minmax_fold(const double *__restrict__ part, int nparts, double *__restrict__ out) {
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int i = threadIdx.x; i < nparts; i += 32) { mn = fmin(mn, part[2 * i]); mx = fmax(mx, part[2 * i + 1]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
}

// ---- joint histogram ------------------------------------------------------------------------
// np.linspace(lo, hi, bins + 1): step = (hi - lo) / bins; edge i = i * step + lo (two roundings), last = hi;
// np.histogramdd widens a degenerate range (lo == hi) by 0.5 on both sides first
__device__ __forceinline__ void make_edges(double lo, double hi, int bins, double *e) {
    if (lo == hi) { lo = lo - 0.5; hi = hi + 0.5; }
    const double step = (hi - lo) / (double)bins;
    for (int i = threadIdx.x; i <= bins; i += blockDim.x) e[i] = i == bins ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
}
__device__ __forceinline__ int find_bin(double v, const double *e, int bins, double inv) {
    int k = (int)((v - e[0]) * inv);
    k = max(0, min(k, bins - 1));
    while (k > 0 && v < e[k]) --k;
    while (k < bins - 1 && v >= e[k + 1]) ++k;
    return k;
}

// dynamic shared memory: bins * bins unsigned counters, then 2 * (bins + 1) doubles of edges.
// hist[ia * bins + ib] counts voxels with a in bin ia and b in bin ib (zeroed by the API call).
template <typename T>
__global__ void __launch_bounds__(512)
k_joint_hist(const void *__restrict__ a, const void *__restrict__ b, size_t n, const double *__restrict__ mm_a,
             const double *__restrict__ mm_b, int bins, unsigned long long *__restrict__ hist) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *cnt = reinterpret_cast<unsigned *>(smem);
    double *ea = reinterpret_cast<double *>(smem + (((size_t)bins * bins * sizeof(unsigned) + 15) & ~(size_t)15)), *eb = ea + bins + 1;
    for (int i = threadIdx.x; i < bins * bins; i += blockDim.x) cnt[i] = 0u;
    make_edges(mm_a[0], mm_a[1], bins, ea);
    make_edges(mm_b[0], mm_b[1], bins, eb);
    __syncthreads();
    const double inva = (double)bins / (ea[bins] - ea[0]), invb = (double)bins / (eb[bins] - eb[0]);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double va = ldd<T>(a, i), vb = ldd<T>(b, i);
        if (va != va || vb != vb) continue;               // np.histogramdd drops samples it cannot place
        atomicAdd(&cnt[find_bin(va, ea, bins, inva) * bins + find_bin(vb, eb, bins, invb)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins * bins; i += blockDim.x)
        if (cnt[i]) atomicAdd(&hist[i], (unsigned long long)cnt[i]);
}

// ---- plane sums (detect_zero_padding) ---------------------------------------------------------
// xs[x] = sum_{y,z} im, ys[y] = sum_{x,z} im, zs[z] = sum_{x,y} im   (float64; zeroed by the API call)
template <typename T>
__global__ void __launch_bounds__(256)
k_axis_sums(const void *__restrict__ im, int X, int Y, int Z, double *__restrict__ xs, double *__restrict__ ys, double *__restrict__ zs) {
    // block = one x plane; a thread owns the z columns tid, tid + 256, ... (<= 8 of them: Z <= 2048) and walks y
    constexpr int KZ = 8;
    const int x = blockIdx.x, lane = threadIdx.x & 31;
    double zacc[KZ] = {0, 0, 0, 0, 0, 0, 0, 0}, xsum = 0.0;
    for (int y = 0; y < Y; ++y) {
        const size_t row = ((size_t)x * Y + y) * Z;
        double rs = 0.0;
#pragma unroll
        for (int k = 0; k < KZ; ++k) {
            const int z = threadIdx.x + 256 * k;
            if (z < Z) { const double v = ldd<T>(im, row + z); zacc[k] += v; rs += v; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, o);
        if (lane == 0 && rs != 0.0) atomicAdd(&ys[y], rs);
        xsum += rs;                                       // meaningful in lane 0 of every warp
    }
    if (lane == 0 && xsum != 0.0) atomicAdd(&xs[x], xsum);
#pragma unroll
    for (int k = 0; k < KZ; ++k) {
        const int z = threadIdx.x + 256 * k;
        if (z < Z && zacc[k] != 0.0) atomicAdd(&zs[z], zacc[k]);
    }
}

// ---- overlap sums (eval_reg_on_sc_seg.py:80-93) -------------------------------------------------
// out[0] = sum(m[fx == 1]), out[1] = sum(m[fx == 0]), out[2] = count(fx == 1), out[3] = count(fx == 0), out[4] = sum(m)
template <typename T>
__global__ void __launch_bounds__(256)
k_overlap_partial(const void *__restrict__ fx, const void *__restrict__ m, size_t n, double *__restrict__ part) {
    __shared__ double s[8][5];
    double a[5] = {0, 0, 0, 0, 0};
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256ull) {
        const double f = ldd<T>(fx, i), v = ldd<T>(m, i);
        if (f == 1.0) { a[0] += v; a[2] += 1.0; }
        if (f == 0.0) { a[1] += v; a[3] += 1.0; }
        a[4] += v;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 5; ++k) s[threadIdx.x >> 5][k] = a[k];
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
        part[blockIdx.x * 5 + threadIdx.x] = t;
    }
}
__global__ void k_overlap_fold(const double *__restrict__ part, int nparts, double *__restrict__ out) {
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int i = 0; i < nparts; ++i) t += part[i * 5 + threadIdx.x];
        out[threadIdx.x] = t;
    }
}

constexpr int METRIC_BLOCKS = 148 * 4;

}  // namespace dfm

using namespace dfm;

extern "C" size_t dfm_metrics_workspace_bytes(void) { return (size_t)METRIC_BLOCKS * 5 * sizeof(double); }

extern "C" int dfm_minmax(const void *a, size_t n, int is_f64, double *out2, double *work, void *stream) {
    DFM_REQUIRE(a && out2 && work && n > 0, DFM_EINVAL, "dfm_minmax: null pointer or empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::min<size_t>(METRIC_BLOCKS, (n + 255) / 256);
    if (is_f64) k_minmax<double><<<blocks, 256, 0, st>>>(a, n, work);
    else k_minmax<float><<<blocks, 256, 0, st>>>(a, n, work);
    k_minmax_fold<<<1, 32, 0, st>>>(work, blocks, out2);
    return check_launch("k_minmax");
}

extern "C" int dfm_joint_hist(const void *a, const void *b, size_t n, int is_f64, const double *minmax_a, const double *minmax_b,
                              int bins, unsigned long long *hist, void *stream) {
    DFM_REQUIRE(a && b && minmax_a && minmax_b && hist, DFM_EINVAL, "dfm_joint_hist: null pointer");
    DFM_REQUIRE(bins >= 1 && bins <= 160, DFM_EINVAL, "dfm_joint_hist: bins must be in 1..160 (got %d)", bins);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)bins * bins * sizeof(unsigned long long), st);
    DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "dfm_joint_hist: %s", cudaGetErrorString(e));
    if (n == 0) return DFM_OK;
    const size_t smem = (((size_t)bins * bins * sizeof(unsigned) + 15) & ~(size_t)15) + 2 * (size_t)(bins + 1) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        e = cudaFuncSetAttribute(k_joint_hist<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_joint_hist<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DFM_REQUIRE(e == cudaSuccess, DFM_ECUDA, "k_joint_hist smem attribute: %s", cudaGetErrorString(e));
        configured = smem;
    }
    const int blocks = (int)std::min<size_t>(148 * 2, (n + 511) / 512);
    if (is_f64) k_joint_hist<double><<<blocks, 512, smem, st>>>(a, b, n, minmax_a, minmax_b, bins, hist);
    else k_joint_hist<float><<<blocks, 512, smem, st>>>(a, b, n, minmax_a, minmax_b, bins, hist);
    return check_launch("k_joint_hist");
}

extern "C" int dfm_axis_sums(const void *im, int X, int Y, int Z, int is_f64, double *xs, double *ys, double *zs, void *stream) {
    DFM_REQUIRE(im && xs && ys && zs && X >= 1 && Y >= 1 && Z >= 1, DFM_EINVAL, "dfm_axis_sums: bad argument");
    DFM_REQUIRE(Z <= 2048, DFM_EUNSUPPORTED, "dfm_axis_sums: Z must be <= 2048 (got %d)", Z);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(xs, 0, X * sizeof(double), st);
    cudaMemsetAsync(ys, 0, Y * sizeof(double), st);
    cudaMemsetAsync(zs, 0, Z * sizeof(double), st);
    dim3 grid(X);
    if (is_f64) k_axis_sums<double><<<grid, 256, 0, st>>>(im, X, Y, Z, xs, ys, zs);
    else k_axis_sums<float><<<grid, 256, 0, st>>>(im, X, Y, Z, xs, ys, zs);
    return check_launch("k_axis_sums");
}

extern "C" int dfm_overlap_sums(const void *fx, const void *m, size_t n, int is_f64, double *out5, double *work, void *stream) {
    DFM_REQUIRE(fx && m && out5 && work, DFM_EINVAL, "dfm_overlap_sums: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(METRIC_BLOCKS, (n + 255) / 256));
    if (is_f64) k_overlap_partial<double><<<blocks, 256, 0, st>>>(fx, m, n, work);
    else k_overlap_partial<float><<<blocks, 256, 0, st>>>(fx, m, n, work);
    k_overlap_fold<<<1, 32, 0, st>>>(work, blocks, out5);
    return check_launch("k_overlap");
}
