// TMA (cp.async.bulk.tensor, SASS UTMALDG) plumbing for the NHWC level buffers.
//
// A level buffer [B][H][W][C] fp32 is described to the TMA unit as a rank-4 tensor (C, W, H, B); a box of
// (channels, width, height, 1) elements starting at (c0, x0, y0, b) -- coordinates may be negative or run past the
// edge: out-of-bound elements are ZERO-FILLED by the hardware, which is the padding a 3x3 convolution tile needs --
// lands in shared memory as a dense [height][width][channels] block.  One thread issues the copy; completion is a
// transaction count on an mbarrier.  The tensor map is built on the host per call (the buffers are caller-owned, so
// their addresses are only known then) with cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint: the
// library links against the CUDA runtime only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"

namespace endo {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// NHWC fp32 buffer [B][H][W][C] (C a multiple of 4: strides are multiples of 16 bytes) -> tensor map with a box of
// (box_c, box_w, box_h, 1).  Returns false when the driver entry point is missing or the arguments are rejected.
static inline bool make_nhwc_map(CUtensorMap* map, const float* base, int B, int H, int W, int C, int box_c, int box_w, int box_h) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one thread: box at (c0, x0, y0, b) -> smem_dst (128-byte aligned), completion on `bar` (post expect_tx with the box bytes first)
__device__ __forceinline__ void load_4d(void* smem_dst, const CUtensorMap* map, int c0, int x0, int y0, int b, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            tc::smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(x0), "r"(y0), "r"(b), "r"(tc::smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace endo
