// TMA / mbarrier helpers shared by the shared-memory kernels (sm_100a inline PTX + the driver's
// cuTensorMapEncodeTiled obtained through the runtime, so libcuda is not a link dependency).
#pragma once
#include <cuda.h>

#include "dfm_common.cuh"

namespace dfm {

// ------------------------------- PTX helpers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    // bounded spin: a barrier that never completes traps instead of hanging the GPU
    for (uint32_t it = 0; !mbar_try_wait(bar, phase); ++it)
        if (it > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// 4-D map over a planar [nvol][X][Y][Z] fp32 tensor with box {bz, by, bx, bc}  -- dfm_tma.cu
bool encode_planar_map(CUtensorMap *tmap, const float *base, int nvol, int X, int Y, int Z, int bx, int by, int bz,
                       int bc);
// 5-D map over a channels-last 3-component field [B][X][Y][Z][3] whose rows (3 Z floats, Z a multiple of 32) are split in
// 96-float pieces: dims {96, 3Z/96, Y, X, B}, box {96, 3Z/96, rows, 1, 1} -- a box lands as `rows` contiguous rows of 3 Z floats
bool encode_cl_rows_map(CUtensorMap *tmap, const float *base, int B, int X, int Y, int Z, int rows);
// TMA needs a 16-byte aligned base and 16-byte multiples for every global stride
bool tma_planar_ok(const float *p, int X, int Y, int Z);

}  // namespace dfm
