// Layout conversion between the reference's channels-last tensors and planar tensors.
#include "dfm_common.cuh"

namespace dfm {

// cl [B][N][C] -> planar [B][C][N]; a block stages a tile of TN voxels x C channels in shared
// memory so both the global read and the global write are contiguous.
template <typename T, bool TO_PLANAR>
__global__ void __launch_bounds__(256)
k_transpose(const T *__restrict__ in, T *__restrict__ out, int C, size_t N) {
    extern __shared__ unsigned char smem_raw[];
    T *tile = reinterpret_cast<T *>(smem_raw);
    const int TN = 256;
    const size_t n0 = (size_t)blockIdx.x * TN;
    const int tn = (int)min((size_t)TN, N - n0);
    const size_t b = blockIdx.y;
    const T *ib = in + b * (size_t)C * N;
    T *ob = out + b * (size_t)C * N;
    const int total = tn * C;
    if (TO_PLANAR) {
        // read the contiguous span cl[n0*C .. (n0+tn)*C)
        for (int t = threadIdx.x; t < total; t += blockDim.x) tile[t] = ib[n0 * C + t];
        __syncthreads();
        for (int t = threadIdx.x; t < total; t += blockDim.x) {
            const int c = t / tn, n = t - c * tn;
            ob[(size_t)c * N + n0 + n] = tile[n * C + c];
        }
    } else {
        for (int t = threadIdx.x; t < total; t += blockDim.x) {
            const int c = t / tn, n = t - c * tn;
            tile[n * C + c] = ib[(size_t)c * N + n0 + n];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < total; t += blockDim.x) ob[n0 * C + t] = tile[t];
    }
}

template <bool TO_PLANAR>
static int transpose(const void *in, void *out, int B, int C, size_t N, int elem_size, void *stream) {
    DFM_REQUIRE(B >= 0 && C >= 1 && B <= 65535, DFM_EINVAL, "dfm layout: bad shape B=%d C=%d", B, C);
    if (B == 0 || N == 0) return DFM_OK;
    DFM_REQUIRE(in && out && in != out, DFM_EINVAL, "dfm layout: null or aliased pointer");
    const size_t smem = (size_t)256 * C * elem_size;
    DFM_REQUIRE(smem <= 48 * 1024, DFM_EUNSUPPORTED, "dfm layout: C*elem_size = %d too large", C * elem_size);
    DFM_REQUIRE((N + 255) / 256 < (1ull << 31), DFM_EINVAL, "dfm layout: N too large");
    dim3 grid((unsigned)((N + 255) / 256), B), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem_size) {
        case 1: k_transpose<uint8_t, TO_PLANAR><<<grid, block, smem, st>>>((const uint8_t *)in, (uint8_t *)out, C, N); break;
        case 2: k_transpose<uint16_t, TO_PLANAR><<<grid, block, smem, st>>>((const uint16_t *)in, (uint16_t *)out, C, N); break;
        case 4: k_transpose<uint32_t, TO_PLANAR><<<grid, block, smem, st>>>((const uint32_t *)in, (uint32_t *)out, C, N); break;
        case 8: k_transpose<uint64_t, TO_PLANAR><<<grid, block, smem, st>>>((const uint64_t *)in, (uint64_t *)out, C, N); break;
        default: return fail(DFM_EINVAL, "dfm layout: elem_size %d not in {1,2,4,8}", elem_size);
    }
    return check_launch("dfm layout");
}

template <bool IN_CL>
__global__ void __launch_bounds__(256)
k_scale_copy3(const float *__restrict__ in, float *__restrict__ out, size_t N, float scale, float *absmax) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (n < N) {
        const float *ib = in + (size_t)blockIdx.y * 3 * N;
        float *ob = out + (size_t)blockIdx.y * 3 * N;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = scale * (IN_CL ? ib[n * 3 + c] : ib[c * N + n]);
            ob[c * N + n] = v;
            m = absmax_fold(m, v);
        }
    }
    if (absmax) block_absmax_commit(m, absmax + blockIdx.y);
}

int scale_copy_to_planar(const float *in, float *out, int B, size_t N, float scale, bool in_cl, float *absmax,
                         cudaStream_t st) {
    dim3 grid((unsigned)((N + 255) / 256), B), block(256);
    if (in_cl) k_scale_copy3<true><<<grid, block, 0, st>>>(in, out, N, scale, absmax);
    else k_scale_copy3<false><<<grid, block, 0, st>>>(in, out, N, scale, absmax);
    return check_launch("scale_copy_to_planar");
}

}  // namespace dfm

extern "C" int dfm_cl_to_planar(const void *cl, void *planar, int B, int C, size_t N, int elem_size, void *stream) {
    return dfm::transpose<true>(cl, planar, B, C, N, elem_size, stream);
}
extern "C" int dfm_planar_to_cl(const void *planar, void *cl, int B, int C, size_t N, int elem_size, void *stream) {
    return dfm::transpose<false>(planar, cl, B, C, N, elem_size, stream);
}
