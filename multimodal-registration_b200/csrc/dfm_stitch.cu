// Sub-volume stitching (3d_reg.py:214-259 get_def_field_from_subvol): weighted average of the
// overlapping tile fields with the pyramid weight 1 - max(|x|,|y|,|z|)/(max+1).  The reference
// materialises one full-size float64 array per tile; here each output voxel gathers from the
// tiles that cover it (every tile voxel is read exactly once), in float64 and in tile order, so
// the result is bit-identical to the reference's.
#include "dfm_common.cuh"

namespace dfm {

constexpr int STITCH_MAX_TILES = 1024;

template <typename Tout, bool IN_CL, bool OUT_CL>
__global__ void __launch_bounds__(256)
k_stitch(const float *__restrict__ tiles, const int *__restrict__ mins, Tout *__restrict__ out, int T, int tx,
         int ty, int tz, int X, int Y, int Z, FastDiv zdiv, uint32_t plane_items) {
    __shared__ int s_min[STITCH_MAX_TILES * 3];
    for (int i = threadIdx.x; i < T * 3; i += blockDim.x) s_min[i] = mins[i];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane_items) return;
    const int y = (int)fast_div(p, zdiv), z = (int)(p - (uint32_t)y * zdiv.d), x = blockIdx.y;
    const int hx = tx / 2, hy = ty / 2, hz = tz / 2;
    const double denom = (double)(max(hx, max(hy, hz)) + 1);      // np.max(w_map) + 1
    const size_t nt = (size_t)tx * ty * tz;
    double sum = 0.0;
    for (int t = 0; t < T; ++t) {                                 // sum_weights (+= in tile order)
        const int lx = x - s_min[3 * t], ly = y - s_min[3 * t + 1], lz = z - s_min[3 * t + 2];
        if (lx < 0 || lx >= tx || ly < 0 || ly >= ty || lz < 0 || lz >= tz) continue;
        const int m = max(abs(lx - hx), max(abs(ly - hy), abs(lz - hz)));
        sum += 1.0 - (double)m / denom;
    }
    if (sum == 0.0) sum = 1.0;                                    // :246
    double acc[3] = {0.0, 0.0, 0.0};
    for (int t = 0; t < T; ++t) {
        const int lx = x - s_min[3 * t], ly = y - s_min[3 * t + 1], lz = z - s_min[3 * t + 2];
        if (lx < 0 || lx >= tx || ly < 0 || ly >= ty || lz < 0 || lz >= tz) continue;
        const int m = max(abs(lx - hx), max(abs(ly - hy), abs(lz - hz)));
        const double w = (1.0 - (double)m / denom) / sum;
        const size_t lo = ((size_t)lx * ty + ly) * tz + lz;
        const float *tb = tiles + (size_t)t * 3 * nt;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = (double)(IN_CL ? __ldg(tb + lo * 3 + c) : __ldg(tb + c * nt + lo));
            acc[c] = __dadd_rn(acc[c], __dmul_rn(w, v));
        }
    }
    const size_t N = (size_t)X * Y * Z, vox = ((size_t)x * Y + y) * Z + z;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (OUT_CL) out[vox * 3 + c] = (Tout)acc[c];
        else out[c * N + vox] = (Tout)acc[c];
    }
}

}  // namespace dfm

using namespace dfm;

extern "C" int dfm_stitch_subvol(const float *tiles, const int *mins, void *out, int T, int tx, int ty, int tz,
                                 int X, int Y, int Z, int out_f64, unsigned flags, void *stream) {
    DFM_REQUIRE(T >= 1 && T <= STITCH_MAX_TILES, DFM_EINVAL, "dfm_stitch_subvol: %d tiles (1..%d supported)", T, STITCH_MAX_TILES);
    DFM_REQUIRE(tx >= 2 && ty >= 2 && tz >= 2 && tx % 2 == 0 && ty % 2 == 0 && tz % 2 == 0, DFM_EINVAL,
                "dfm_stitch_subvol: tile shape (%d,%d,%d) must be even (the reference's weight grid is [-s/2, s/2))", tx, ty, tz);
    DFM_REQUIRE(X >= 1 && Y >= 1 && Z >= 1 && X <= 65535 && (uint64_t)X * Y * Z < (1ull << 31), DFM_EINVAL,
                "dfm_stitch_subvol: bad volume shape");
    DFM_REQUIRE((uint64_t)Y * Z * (uint64_t)Z < (1ull << 32), DFM_EINVAL, "dfm_stitch_subvol: Y*Z*Z must be < 2^32");
    DFM_REQUIRE(tiles && mins && out, DFM_EINVAL, "dfm_stitch_subvol: null pointer");
    const uint32_t plane = (uint32_t)Y * Z;
    dim3 grid((plane + 255) / 256, X), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    const FastDiv fd = make_fastdiv(Z);
    const bool icl = flags & DFM_FIELD_IN_CL, ocl = flags & DFM_FIELD_OUT_CL;
#define DFM_GO(TT, I, O) k_stitch<TT, I, O><<<grid, block, 0, st>>>(tiles, mins, (TT *)out, T, tx, ty, tz, X, Y, Z, fd, plane)
#define DFM_GO2(TT) do { if (icl) { if (ocl) DFM_GO(TT, true, true); else DFM_GO(TT, true, false); } \
                         else { if (ocl) DFM_GO(TT, false, true); else DFM_GO(TT, false, false); } } while (0)
    if (out_f64) DFM_GO2(double); else DFM_GO2(float);
#undef DFM_GO2
#undef DFM_GO
    return check_launch("dfm_stitch_subvol");
}
