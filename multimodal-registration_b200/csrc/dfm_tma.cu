// Host side of the TMA helpers: tensor-map encoding through the driver entry point.
#include "dfm_tma.cuh"

namespace dfm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

bool encode_planar_map(CUtensorMap *tmap, const float *base, int nvol, int X, int Y, int Z, int bx, int by, int bz,
                       int bc) {
    if (!encode_fn()) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)nvol};
    const cuuint64_t strides[3] = {(cuuint64_t)Z * 4, (cuuint64_t)Y * Z * 4, (cuuint64_t)X * Y * Z * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, (cuuint32_t)bc};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode_fn()(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool encode_cl_rows_map(CUtensorMap *tmap, const float *base, int B, int X, int Y, int Z, int rows) {
    if (!encode_fn() || Z % 32 != 0) return false;
    const cuuint64_t row = (cuuint64_t)Z * 3 * 4;
    const cuuint64_t dims[5] = {96, (cuuint64_t)(3 * Z / 96), (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)B};
    const cuuint64_t strides[4] = {96 * 4, row, row * (cuuint64_t)Y, row * (cuuint64_t)Y * (cuuint64_t)X};
    const cuuint32_t box[5] = {96, (cuuint32_t)(3 * Z / 96), (cuuint32_t)rows, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return encode_fn()(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void *)base, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tma_planar_ok(const float *p, int X, int Y, int Z) {
    (void)X; (void)Y;
    return encode_fn() != nullptr && Z % 4 == 0 && aligned16(p);
}

}  // namespace dfm
