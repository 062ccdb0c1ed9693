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

// =====================================================================================================
// Sparse SfM rasteriser (reference utils.get_torch_training_data, utils.py:460-612): the step BEFORE the hot path.
// Projects the M (a few hundred) SfM points of a sequence into the two views of a pair and scatters the sparse depth,
// depth-mask, flow and flow-mask images the loss stack consumes; the reference does it with numpy on the CPU per sample
// and the DataLoader then copies ten H x W images to the GPU (train.py:255-270).  Here the images are produced on the
// device.  Semantics kept exactly: float64 projection, np.round (half to even), `a[idx] = v` with duplicate indices =
// the LAST point (in point order) wins, flow normalised in float32, |flow| > 5 outliers cleared (utils.py:567-574).
// =====================================================================================================
namespace endo {

struct RasterArgs {
    const double* pts;            // [M][4] homogeneous points
    const double* P;              // [2][3][4] projection matrices of the two views
    const double* E;              // [2][4][4] extrinsic matrices
    const float* vis;             // [2][M] visibility of each point in each view (> 0.5 = visible)
    const float* clean;           // [M] inlier flags or nullptr (utils.py:505-508)
    const unsigned char* mask;    // [H][W] boundary mask, 255 = inside
    int M, H, W;
    int* winner;                  // [2][H*W], -1 = empty
    double* uvz;                  // [2][M][3] rounded pixel position (u, v) and camera-space depth
    float* depth_mask; float* depth; float* flow_mask; float* flow;    // [2][H][W][1|2] like the reference's return values
};

__global__ void raster_project_kernel(const RasterArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * A.M) return;
    const int view = i / A.M, m = i - view * A.M;
    const double* X = A.pts + (size_t)m * 4;
    const double* P = A.P + view * 12;
    const double* E = A.E + view * 16;
    double p[3], c[4];
#pragma unroll
    for (int r = 0; r < 3; ++r) p[r] = ((P[r * 4] * X[0] + P[r * 4 + 1] * X[1]) + P[r * 4 + 2] * X[2]) + P[r * 4 + 3] * X[3];   // einsum order (:483)
#pragma unroll
    for (int r = 0; r < 4; ++r) c[r] = ((E[r * 4] * X[0] + E[r * 4 + 1] * X[1]) + E[r * 4 + 2] * X[2]) + E[r * 4 + 3] * X[3];
    const double u = rint(p[0] / p[2]), v = rint(p[1] / p[2]);      // np.round: half to even (:484)
    const double z = c[2] / c[3];                                   // :486
    double* o = A.uvz + ((size_t)view * A.M + m) * 3;
    o[0] = u; o[1] = v; o[2] = z;
    const bool visible = A.vis[(size_t)view * A.M + m] > 0.5f && (!A.clean || A.clean[m] > 0.5f);      // :503-515
    if (visible && u <= (double)(A.W - 1) && u >= 0.0 && v <= (double)(A.H - 1) && v >= 0.0 && z > 0.0) {   // :521-525
        const int loc = (int)(u + v * (double)A.W);                 // :527-529
        if (A.mask[loc] == 255) atomicMax(A.winner + (size_t)view * A.H * A.W + loc, m);   // :530-533; last point wins
    }
}

__global__ void raster_write_kernel(const RasterArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * A.M) return;
    const int view = i / A.M, m = i - view * A.M;
    const double* mine = A.uvz + ((size_t)view * A.M + m) * 3;
    const double* other = A.uvz + ((size_t)(1 - view) * A.M + m) * 3;
    const double u = mine[0], v = mine[1], z = mine[2];
    const bool visible = A.vis[(size_t)view * A.M + m] > 0.5f && (!A.clean || A.clean[m] > 0.5f);
    if (!(visible && u <= (double)(A.W - 1) && u >= 0.0 && v <= (double)(A.H - 1) && v >= 0.0 && z > 0.0)) return;
    const int loc = (int)(u + v * (double)A.W);
    const size_t img = (size_t)view * A.H * A.W;
    if (A.mask[loc] != 255 || A.winner[img + loc] != m) return;
    // flow = position in the other view - position in this view (:548-557), stored as float32, then normalised (:559-562)
    float fx = (float)(other[0] - u), fy = (float)(other[1] - v);
    fx = fx / (float)A.W; fy = fy / (float)A.H;
    const bool outlier = fabsf(fx) > 5.0f || fabsf(fy) > 5.0f;      // :564-574
    A.flow[(img + loc) * 2] = outlier ? 0.f : fx;
    A.flow[(img + loc) * 2 + 1] = outlier ? 0.f : fy;
    A.flow_mask[img + loc] = outlier ? 0.f : 1.f;
    A.depth[img + loc] = (float)z;                                  // :585-590
    A.depth_mask[img + loc] = 1.f;
}

}  // namespace endo

extern "C" size_t endo_rasterize_workspace_bytes(int M, int H, int W) {
    if (M < 0 || H <= 0 || W <= 0) return 0;
    return sizeof(int) * 2 * (size_t)H * W + sizeof(double) * 6 * (size_t)(M > 0 ? M : 1) + 64;
}

extern "C" int endo_rasterize_pair(const double* points, const double* projections, const double* extrinsics, const float* visibility,
                                   const float* clean, const unsigned char* mask_boundary, int M, int H, int W, float* depth_mask,
                                   float* depth, float* flow_mask, float* flow, void* ws, size_t ws_bytes, endo_stream_t stream) {
    if (M < 0 || H <= 0 || W <= 0) return ENDO_ERR_BAD_SHAPE;
    if (!projections || !extrinsics || !mask_boundary || !depth_mask || !depth || !flow_mask || !flow || (M > 0 && (!points || !visibility)))
        return ENDO_ERR_BAD_POINTER;
    if (!ws || ws_bytes < endo_rasterize_workspace_bytes(M, H, W) || (reinterpret_cast<uintptr_t>(ws) & 7u)) return ENDO_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t hw = (size_t)H * W;
    endo::RasterArgs A{points, projections, extrinsics, visibility, clean, mask_boundary, M, H, W,
                       nullptr, reinterpret_cast<double*>(ws), depth_mask, depth, flow_mask, flow};
    A.winner = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + sizeof(double) * 6 * (size_t)(M > 0 ? M : 1));
    ENDO_CUDA(cudaMemsetAsync(A.winner, 0xff, sizeof(int) * 2 * hw, s));
    ENDO_CUDA(cudaMemsetAsync(depth_mask, 0, sizeof(float) * 2 * hw, s));
    ENDO_CUDA(cudaMemsetAsync(depth, 0, sizeof(float) * 2 * hw, s));
    ENDO_CUDA(cudaMemsetAsync(flow_mask, 0, sizeof(float) * 2 * hw, s));
    ENDO_CUDA(cudaMemsetAsync(flow, 0, sizeof(float) * 4 * hw, s));
    if (M == 0) return ENDO_OK;
    endo::raster_project_kernel<<<endo::cdiv(2 * M, 256), 256, 0, s>>>(A);
    ENDO_CHECK_LAUNCH();
    endo::raster_write_kernel<<<endo::cdiv(2 * M, 256), 256, 0, s>>>(A);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

// =====================================================================================================
// Image side of the input pipeline ("next" row N3): utils.get_pair_color_imgs (utils.py:441-457) after the JPEG decode --
// cv2.resize(fx = fy = 1/downsampling, INTER_LINEAR, 8-bit) -> crop -> BGR2RGB -- and the normalisation dataset.py:148,446-453
// applies (albumentations Normalize(0.5, 0.5, 255) + img_to_tensor).  The reference does this per sample on DataLoader workers
// from 1080x1920 frames; here one thread produces one output pixel straight from the decoded frame in HBM, bit for bit like
// cv::resize's fixed-point bilinear filter (11-bit coefficients, int horizontal pass, truncating vertical pass; see
// oracle/pipeline.py for the restatement and its pinning against cv2 itself).  This file is compiled with -fmad=false: the
// source coordinate (d + 0.5) * scale - 0.5 must round like the two separate double operations cv::resize performs.
// =====================================================================================================
namespace endo {

struct ResizeArgs {
    const unsigned char* src; int sh, sw;            // decoded frame, HWC, 3 channels (BGR as cv2.imread returns it)
    double scale_x, scale_y;                         // source pixels per destination pixel (= downsampling factor)
    int start_h, start_w, H, W;                      // crop origin inside the resized image, output size
    int swap_rb;                                     // 1: BGR -> RGB
    unsigned char* out_u8;                           // [H][W][3] or nullptr
    float* out_norm;                                 // [3][H][W] float32, (v - 127.5) * (1 / 127.5), or nullptr
};

__global__ void __launch_bounds__(256) resize_crop_kernel(const ResizeArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.H * A.W) return;
    const int y = i / A.W, x = i - y * A.W;
    const int dx = x + A.start_w, dy = y + A.start_h;
    // horizontal: fractional weight AND position clamp at the borders
    float fx = (float)(((double)dx + 0.5) * A.scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= A.sw - 1) { fx = 0.f; sx = A.sw - 1; }
    const int sx1 = min(sx + 1, A.sw - 1);
    const int a1 = __float2int_rn(fx * 2048.f), a0 = __float2int_rn((1.f - fx) * 2048.f);
    // vertical: only the row indices clamp, the weights stay
    float fy = (float)(((double)dy + 0.5) * A.scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy -= (float)sy;
    const int b1 = __float2int_rn(fy * 2048.f), b0 = __float2int_rn((1.f - fy) * 2048.f);
    const int y0 = min(max(sy, 0), A.sh - 1), y1 = min(max(sy + 1, 0), A.sh - 1);
    const unsigned char* r0 = A.src + (size_t)y0 * A.sw * 3;
    const unsigned char* r1 = A.src + (size_t)y1 * A.sw * 3;
    const float mean = 0.5f * 255.0f, den = 1.0f / (0.5f * 255.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
        const int s1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
        const int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        const int oc = A.swap_rb ? 2 - c : c;
        if (A.out_u8) A.out_u8[(size_t)i * 3 + oc] = (unsigned char)v;
        if (A.out_norm) A.out_norm[(size_t)oc * A.H * A.W + i] = ((float)v - mean) * den;
    }
}

}  // namespace endo

extern "C" int endo_resize_crop_u8(const unsigned char* src_bgr, int src_h, int src_w, double downsampling, int start_h, int end_h,
                                   int start_w, int end_w, int swap_rb, unsigned char* out_u8, float* out_norm, endo_stream_t stream) {
    if (src_h <= 0 || src_w <= 0 || !(downsampling > 0.0)) return ENDO_ERR_BAD_SHAPE;
    if (!src_bgr || (!out_u8 && !out_norm)) return ENDO_ERR_BAD_POINTER;
    const double f = 1.0 / downsampling;                     // fx = fy of cv2.resize(img, (0, 0), fx, fy)
    const int dh = (int)nearbyint(src_h * f), dw = (int)nearbyint(src_w * f);       // saturate_cast<int>: round half to even
    if (start_h < 0 || start_w < 0 || end_h <= start_h || end_w <= start_w || end_h > dh || end_w > dw) return ENDO_ERR_BAD_SHAPE;
    // cv::resize switches to its INTER_AREA fast path at a factor of exactly 2; it equals the bilinear formula except on the
    // last row / column of odd-sized sources
    if (fabs(1.0 / f - 2.0) < 2.220446049250313e-16 && ((src_h | src_w) & 1)) return ENDO_ERR_CONFIG;
    endo::ResizeArgs A{src_bgr, src_h, src_w, 1.0 / f, 1.0 / f, start_h, start_w, end_h - start_h, end_w - start_w, swap_rb, out_u8, out_norm};
    const long long n = (long long)A.H * A.W;
    endo::resize_crop_kernel<<<endo::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(A);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}
