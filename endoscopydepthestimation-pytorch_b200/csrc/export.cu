// Depth map -> coloured point cloud (reference utils.py:825-852, called by evaluate.py:337-341): the step AFTER the network
// in the evaluation / export path.  The reference walks the H x W image in a pure-Python double loop; here one pass counts the
// surviving pixels per 1,024-pixel block and a second pass writes them COMPACTED IN ROW-MAJOR ORDER (the order the reference
// appends them in), so the result is identical element for element.  fp32 arithmetic in the reference's operation order
// ((w - cx) / fx * z; this file is compiled with -fmad=false).
#include "common.cuh"

namespace endo {

constexpr int kPcThreads = 1024;

struct PcArgs {
    const float* depth; const unsigned char* color; const float* mask;
    float fx, fy, cx, cy;
    int H, W, ds, use_thr;
    float min_thr, max_thr;
    float* points; int* count; int* block_counts;
};

__device__ __forceinline__ bool pc_keep(const PcArgs& A, int p, int& h, int& w) {
    h = p / A.W; w = p - h * A.W;
    if (p >= A.H * A.W) return false;
    if ((h % A.ds) != 0 || (w % A.ds) != 0 || !(A.mask[p] > 0.5f)) return false;
    if (A.use_thr) {
        const unsigned char b = A.color[3 * p], g = A.color[3 * p + 1], r = A.color[3 * p + 2];
        const float mx = (float)max(max(r, g), b), mn = (float)min(min(r, g), b);
        if (!(mx >= A.max_thr && mn <= A.min_thr)) return false;
    }
    return true;
}

__global__ void __launch_bounds__(kPcThreads)
pc_count_kernel(const PcArgs A) {
    __shared__ int warp_cnt[kPcThreads / 32];
    const int p = blockIdx.x * kPcThreads + threadIdx.x;
    int h, w;
    const unsigned bal = __ballot_sync(0xffffffffu, pc_keep(A, p, h, w));
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < kPcThreads / 32; ++i) s += warp_cnt[i];
        A.block_counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kPcThreads)
pc_write_kernel(const PcArgs A) {
    __shared__ int warp_off[kPcThreads / 32 + 1];
    __shared__ int base_s;
    const int p = blockIdx.x * kPcThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int h, w;
    const bool keep = pc_keep(A, p, h, w);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_off[warp + 1] = __popc(bal);
    // exclusive prefix of the block counts before this block (a few hundred blocks: one strided pass + tree)
    int part = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += kPcThreads) part += A.block_counts[i];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ int red[kPcThreads / 32];
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < kPcThreads / 32; ++i) s += red[i];
        base_s = s;
        warp_off[0] = 0;
        for (int i = 1; i <= kPcThreads / 32; ++i) warp_off[i] += warp_off[i - 1];
        if (blockIdx.x == gridDim.x - 1) A.count[0] = s + warp_off[kPcThreads / 32];
    }
    __syncthreads();
    if (keep) {
        const int idx = base_s + warp_off[warp] + __popc(bal & ((1u << lane) - 1u));
        const float z = A.depth[p];
        const float x = ((float)w - A.cx) / A.fx * z;            // utils.py:840-841
        const float y = ((float)h - A.cy) / A.fy * z;
        float* o = A.points + (size_t)idx * 6;
        o[0] = x; o[1] = y; o[2] = z;
        o[3] = (float)A.color[3 * p + 2]; o[4] = (float)A.color[3 * p + 1]; o[5] = (float)A.color[3 * p];   // (x, y, z, r, g, b), :842-849
    }
}

}  // namespace endo

using namespace endo;

extern "C" size_t endo_point_cloud_workspace_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return sizeof(int) * (size_t)cdiv((long long)H * W, kPcThreads) + 64;
}

extern "C" int endo_point_cloud_from_depth(const float* depth, const unsigned char* color_bgr, const float* mask, float fx, float fy,
                                           float cx, float cy, int H, int W, int downsampling, int use_threshold, float min_threshold,
                                           float max_threshold, float* points, int* count, void* ws, size_t ws_bytes,
                                           endo_stream_t stream) {
    if (H <= 0 || W <= 0 || downsampling <= 0) return ENDO_ERR_BAD_SHAPE;
    if (!depth || !color_bgr || !mask || !points || !count) return ENDO_ERR_BAD_POINTER;
    if (!ws || ws_bytes < endo_point_cloud_workspace_bytes(H, W)) return ENDO_ERR_WORKSPACE;
    PcArgs A{depth, color_bgr, mask, fx, fy, cx, cy, H, W, downsampling, use_threshold, min_threshold, max_threshold, points, count,
             reinterpret_cast<int*>(ws)};
    const int nb = cdiv((long long)H * W, kPcThreads);
    cudaStream_t s = (cudaStream_t)stream;
    pc_count_kernel<<<nb, kPcThreads, 0, s>>>(A);
    ENDO_CHECK_LAUNCH();
    pc_write_kernel<<<nb, kPcThreads, 0, s>>>(A);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}
