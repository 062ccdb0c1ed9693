/*
 * endo_b200.h -- C ABI of libendo_b200.so (hand-written CUDA for sm_100a).
 *
 * Drop-in boundary for the training hot path of lppllppl920/EndoscopyDepthEstimation-Pytorch
 * (reference train.py:272-328).  The reference is pure Python/PyTorch and has no FFI of its own;
 * each entry point below replaces the body of one reference nn.Module.forward (and the autograd
 * backward PyTorch derives for it) and is what a ctypes binding inside the reference's models.py /
 * losses.py would call.  The reference interface each function replaces is
 * cited as file:line into /root/reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 data unless stated; images are NCHW with
 *     C = 1 (or 2 for flow, 3 for colours); pose inputs are row-major t[B*3], R[B*9], K[B*9];
 *   - the caller owns all memory (PyTorch's caching allocator); the library never allocates, frees or
 *     keeps a pointer after the call, and never synchronises the device;
 *   - the CURRENT device of the calling thread must be the device that owns `stream` and the buffers (per-device kernel
 *     attributes are configured lazily for the current device);
 *   - all work is ordered with respect to `stream` (a cudaStream_t): when a call returns, everything it enqueued
 *     happens-before whatever the caller enqueues on `stream` next.  Functions are re-entrant and keep no per-call
 *     state.  One exception to "only on `stream`" is documented at endo_net_bwd: the weight-gradient kernels run on a
 *     library-owned side stream forked from / joined into `stream` with events (capturable in a CUDA graph after one
 *     warm-up call; switched off with ENDO_NET_SINGLE_STREAM);
 *   - `ws` is scratch of at least endo_*_workspace_bytes(...) bytes, 16-byte aligned, whose first
 *     ENDO_WS_HEADER_BYTES bytes must be zero on entry (they hold self-resetting arrival counters and
 *     are zero again when the enqueued work has finished).  One `ws` must not be shared by calls that
 *     may run concurrently on different streams;
 *   - return value: 0 = ENDO_OK, otherwise an ENDO_ERR_* code (endo_strerror()).  No C++ exception
 *     crosses the boundary.  NaN/Inf in the data propagate as values (train.py:317 relies on that).
 */
#ifndef ENDO_B200_H
#define ENDO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* endo_stream_t; /* cudaStream_t */

enum {
    ENDO_OK = 0,
    ENDO_ERR_BAD_SHAPE = 1,   /* non-positive dims, H/W not a multiple of what the op needs */
    ENDO_ERR_BAD_POINTER = 2, /* NULL or misaligned pointer */
    ENDO_ERR_WORKSPACE = 3,   /* workspace missing or too small */
    ENDO_ERR_CUDA = 4,        /* a CUDA runtime call / kernel launch failed (cudaPeekAtLastError) */
    ENDO_ERR_CONFIG = 5,      /* unsupported network configuration */
    ENDO_ERR_NO_DEVICE = 6    /* not running on an sm_100 device */
};
#define ENDO_WS_HEADER_BYTES 256

int endo_version(void);                 /* 100 * major + minor */
const char* endo_strerror(int code);
/* number of kernels this library has launched in the calling process (monotonic; for bench.py's gpu_launches) */
unsigned long long endo_launch_count(void);
/* Per-category device timing for bench.py's roofline: while enabled, each entry point brackets its launches
 * with CUDA events on the launching stream; endo_prof_collect() synchronises the device, ADDS the elapsed
 * milliseconds / launch-site counts per category to ms[] / counts[] (endo_prof_categories() entries) and
 * clears the records.  Categories that enqueue several kernels back to back (depth_scale, optimizer) are
 * timed as one span. */
void endo_prof_enable(int on);
int endo_prof_categories(void);
const char* endo_prof_category_name(int category);
int endo_prof_collect(double* ms, unsigned long long* counts);

/* ------------------------------------------------------------------------------------------------
 * DepthScalingLayer.forward  (models.py:346-363)   x = [depth, sparse_depth, sparse_mask]
 *   scaled[B,1,H,W] = s_b * depth;  *norm_std = mean_{i,j}(std_j / s_i)  (the reference's broadcast
 *   of a [B] tensor against a [B,1,1,1] tensor, models.py:363);  stats[B*4] = {s, sum(am), std, mean_sd}
 *   is kept by the caller for the backward.
 * backward: g_depth = s*g + d s/d depth * sum(g*depth)  (gradient w.r.t. depth of the first output;
 *   the second output is a monitoring scalar that train.py never differentiates).
 * ---------------------------------------------------------------------------------------------- */
size_t endo_depth_scale_workspace_bytes(int B, int H, int W);
int endo_depth_scale_fwd(const float* depth, const float* sparse_depth, const float* sparse_mask,
                         float* scaled, float* norm_std, float* stats, int B, int H, int W, float eps,
                         void* ws, size_t ws_bytes, endo_stream_t stream);
int endo_depth_scale_bwd(const float* g_scaled, const float* depth, const float* sparse_depth,
                         const float* stats, float* g_depth, int B, int H, int W, float eps,
                         void* ws, size_t ws_bytes, endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * FlowfromDepthLayer.forward  (models.py:370-374 -> :433-451 -> :377-429)
 *   x = [depth, mask, t, R, K]  ->  flow[B,2,H,W]
 * ---------------------------------------------------------------------------------------------- */
int endo_flow_from_depth_fwd(const float* depth, const float* mask, const float* t, const float* R,
                             const float* K, float* flow, int B, int H, int W, endo_stream_t stream);
int endo_flow_from_depth_bwd(const float* g_flow, const float* depth, const float* mask, const float* t,
                             const float* R, const float* K, float* g_depth, int B, int H, int W,
                             endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * DepthWarpingLayer.forward  (models.py:460-465 -> _depth_warping :469-554, _bilinear_interpolate :325-336)
 *   x = [depth_1, depth_2, mask, t, R, K] -> warped[B,1,H,W], intersect[B,1,H,W] in {0,1}
 * backward: g_d1 (dense) and g_d2 (4-tap scatter-add; the function zeroes g_d2 first).
 * ---------------------------------------------------------------------------------------------- */
int endo_depth_warp_fwd(const float* d1, const float* d2, const float* mask, const float* t,
                        const float* R, const float* K, float* warped, float* intersect, int B, int H,
                        int W, float eps, endo_stream_t stream);
int endo_depth_warp_bwd(const float* g_warped, const float* d1, const float* d2, const float* mask,
                        const float* t, const float* R, const float* K, float* g_d1, float* g_d2, int B,
                        int H, int W, float eps, endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Losses.  Every *_fwd writes the scalar loss to loss[0] and per-sample sums to stats (kept for the
 * backward); every *_bwd takes g_loss as a DEVICE scalar (upstream gradient) so no host sync is needed.
 *
 * SparseMaskedL1Loss.forward (losses.py:62-66)  x = [flows[B,2,H,W], flows_from_depth[B,2,H,W], masks[B,1,H,W]]
 *   stats[B*2] = {sum m|f-f^|, sum m};  bwd writes g wrt flows_from_depth (and g wrt flows if g_flows != NULL)
 * NormalizedDistanceLoss.forward (losses.py:122-146)  x = [depth, warped, intersect, K]
 *   stats[B*4] = {numerator sum, denominator, mean_value, sum m};  bwd writes g_depth and g_warped
 * ScaleInvariantLoss.forward (losses.py:22-32)  x = [predicted, goal, boundaries]
 *   stats[B*4] = {sum r^2, sum r, sum b, 0};  bwd writes g_pred (and g_goal if != NULL)
 * ---------------------------------------------------------------------------------------------- */
size_t endo_loss_workspace_bytes(int B, int H, int W);
int endo_sparse_l1_fwd(const float* flows, const float* flows_from_depth, const float* masks, float* loss,
                       float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                       endo_stream_t stream);
int endo_sparse_l1_bwd(const float* g_loss, const float* flows, const float* flows_from_depth,
                       const float* masks, const float* stats, float* g_flows_from_depth, float* g_flows,
                       int B, int H, int W, float eps, endo_stream_t stream);
int endo_norm_dist_fwd(const float* depth, const float* warped, const float* intersect, const float* K,
                       float* loss, float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                       endo_stream_t stream);
int endo_norm_dist_bwd(const float* g_loss, const float* depth, const float* warped, const float* intersect,
                       const float* K, const float* stats, float* g_depth, float* g_warped, int B, int H,
                       int W, float eps, endo_stream_t stream);
int endo_scale_inv_fwd(const float* pred, const float* goal, const float* boundaries, float* loss,
                       float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                       endo_stream_t stream);
int endo_scale_inv_bwd(const float* g_loss, const float* pred, const float* goal, const float* boundaries,
                       const float* stats, float* g_pred, float* g_goal, int B, int H, int W, float eps,
                       endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * FCDenseNet.forward / autograd backward  (models.py:100-187; FCDenseNet57 = models.py:190-194)
 *
 * endo_net_config mirrors FCDenseNet.__init__'s arguments (models.py:101-103).  Parameters and BN
 * buffers are passed as ONE flat fp32 array each, laid out in state_dict() order with every tensor
 * in its PyTorch logical layout (conv weight OIHW); endo_net_param_count / endo_net_buffer_count give
 * the sizes (num_batches_tracked is kept by the host wrapper, it is an int64 counter).
 *
 * fwd: x[B,3,H,W] (NCHW) -> y[B,1,H,W] = |finalConv(...)|.  `groups` splits the batch into that many
 *   consecutive sub-batches with independent BatchNorm statistics (groups = 2 runs the two
 *   net(colors_1), net(colors_2) calls of train.py:276-277 as one launch sequence; the running
 *   buffers then receive two successive momentum updates, in order).  training = 0 uses the running
 *   statistics (evaluate.py).  `acts` (endo_net_activation_bytes) receives every activation the
 *   backward needs and must stay untouched until endo_net_bwd has run.
 * bwd: g_y[B,1,H,W] -> g_params (flat, same layout as params; ACCUMULATED into if accumulate != 0,
 *   else overwritten) and optionally g_x.  `scratch` (endo_net_backward_scratch_bytes) is transient.
 *   Unless ENDO_NET_SINGLE_STREAM is set, the weight-gradient kernels (which only feed g_params) are enqueued on a side
 *   stream borrowed from a per-device pool for the duration of the call: forked from `stream` after the kernels they
 *   depend on, joined back into `stream` before the function returns ON EVERY EXIT PATH (also on errors), so the
 *   caller sees plain stream semantics.  Concurrent calls borrow different side streams.
 * math: ENDO_MATH_FP32 = fp32 FFMA (parity path, matches the reference's CPU fp32 results);
 *       ENDO_MATH_TF32 = tcgen05 tensor-core tiles, tf32 operands (what cuDNN runs the reference's convolutions in
 *       by default), fp32 accumulation in TMEM; ENDO_MATH_TF32X3 = tcgen05 with error-compensated operands in the
 *       forward (x = hi + lo, three tf32 MMAs per product: fp32-grade depth maps and losses on the tensor cores);
 *       gradients as in ENDO_MATH_TF32 (tf32 data gradient, bf16 weight gradient, fp32 accumulation);
 *       ENDO_MATH_BF16 = tcgen05 with activations and weights of the 3x3 convolutions rounded to bf16 (one kind::f16 MMA per
 *       product, fp32 accumulation in TMEM, fp32 BatchNorm statistics, fp32 master weights; the 1x1 TransitionDown GEMMs use
 *       tf32 operands): BASELINE.json configs[2]; gradients as in ENDO_MATH_TF32;
 *       ENDO_MATH_BF16X3 = the same with two-term bf16 operands (x = b1 + b2, 16 significant bits, three kind::f16 MMAs
 *       of K = 16 per product: half the MMA count of 3xTF32, depth maps within ~2e-5 of fp32).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int in_channels;
    int n_down;               /* len(down_blocks) == len(up_blocks), <= 8 */
    int down_layers[8];
    int up_layers[8];
    int bottleneck_layers;
    int growth_rate;
    int first_conv_channels;
    int n_classes;            /* only 1 is supported */
} endo_net_config;

enum { ENDO_MATH_FP32 = 0, ENDO_MATH_TF32 = 1, ENDO_MATH_BF16 = 2, ENDO_MATH_TF32X3 = 3, ENDO_MATH_BF16X3 = 4 };
/* flags OR-ed into the `math` argument of endo_net_fwd / endo_net_bwd (per call, no global state) */
enum {
    ENDO_MATH_MASK = 0xFF,
    ENDO_NET_SINGLE_STREAM = 0x100  /* endo_net_bwd: enqueue the weight-gradient kernels on `stream` too (no side stream) */
};

long long endo_net_param_count(const endo_net_config* cfg);
long long endo_net_buffer_count(const endo_net_config* cfg); /* running_mean + running_var floats */
size_t endo_net_activation_bytes(const endo_net_config* cfg, int B, int H, int W);
size_t endo_net_backward_scratch_bytes(const endo_net_config* cfg, int B, int H, int W);
int endo_net_fwd(const endo_net_config* cfg, const float* x, const float* params, float* bn_buffers,
                 float* y, void* acts, size_t acts_bytes, int B, int H, int W, int groups, int training,
                 int math, endo_stream_t stream);
int endo_net_bwd(const endo_net_config* cfg, const float* g_y, const float* x, const float* params,
                 float* g_params, float* g_x, void* acts, size_t acts_bytes, void* scratch,
                 size_t scratch_bytes, int B, int H, int W, int groups, int accumulate, int math,
                 endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser tail of train.py:327-328: clip_grad_norm_(params, max_norm) followed by
 * SGD(momentum, no weight decay, no nesterov).step() on the flat parameter / gradient / momentum
 * arrays.  first_step != 0 initialises the momentum buffer with the (clipped) gradient like
 * torch.optim.SGD does.  grad_norm_out[0] receives the pre-clip total norm.  If `finite_flag` is
 * non-NULL and finite_flag[0] == 0 the update is skipped (device-side version of the NaN guard,
 * train.py:317-322).
 * ---------------------------------------------------------------------------------------------- */
size_t endo_sgd_workspace_bytes(long long n);
int endo_sgd_clip_step(float* params, float* grads, float* momentum_buf, long long n, float lr,
                       float momentum, float max_norm, int first_step, const float* finite_flag,
                       float* grad_norm_out, void* ws, size_t ws_bytes, endo_stream_t stream);

/* The same step with NOTHING step-dependent in the launch parameters, so that a CUDA graph of the whole optimisation
 * step (train.py:272-328) can be replayed under a learning-rate schedule (train.py:203, scheduler.py): the learning rate is
 * read from the device scalar lr_dev[0], and `momentum_buf` must be ZERO before the first step (momentum * 0 + g == g is
 * exactly torch.optim.SGD's first-step initialisation, so no first_step flag is needed). */
int endo_sgd_clip_step_dev(float* params, float* grads, float* momentum_buf, long long n, const float* lr_dev,
                           float momentum, float max_norm, const float* finite_flag, float* grad_norm_out, void* ws,
                           size_t ws_bytes, endo_stream_t stream);

/* Development aid: copies the clock64() trace that CTA 0 of the last tensor-core forward launch recorded when
 * ENDO_TC_DEBUG has bit 4 set (tools/trace_fwd.py decodes it).  Synchronises the device. */
int endo_debug_trace_read(long long* host_out, int n);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 bring-up probe (tests only): D[128][N] = A[shift..shift+128][K] * B[N][K]^T computed with
 * tcgen05.mma from operands staged in the shared-memory layouts the convolution kernels use.
 * fmt: 2 = tf32, 1 = bf16; a_mn_major / b_mn_major select the MN-major canonical layout.
 * swizzle: 0 = SWIZZLE_NONE, 2 = SWIZZLE_128B (K-major, K*elem a multiple of 128 B; row sliding uses the
 * descriptor's base_offset).  reps > 1 repeats the MMA chain (accumulating) and *cycles receives the SM clock
 * cycles from first issue to completion: the per-MMA cost of an operand layout can be read off directly.
 * ---------------------------------------------------------------------------------------------- */
int endo_tc_probe(const float* A, const float* B, float* D, int a_rows, int N, int K, int shift, int fmt,
                  int a_mn_major, int b_mn_major, int swizzle, int reps, long long* cycles, int rotate,
                  endo_stream_t stream);   /* rotate > 1 (timing only): reps cycle over that many accumulator tiles */

/* ------------------------------------------------------------------------------------------------
 * utils.point_cloud_from_depth (utils.py:825-852; evaluate.py:337-341): depth map [H,W] + BGR uint8 image [H,W,3] + mask [H,W]
 * -> points[N][6] = (x, y, z, r, g, b) of the pixels with h % downsampling == 0, w % downsampling == 0, mask > 0.5 (and, if
 * use_threshold, max(r,g,b) >= max_threshold and min(r,g,b) <= min_threshold), in ROW-MAJOR order like the reference's loop;
 * count[0] = N.  `points` must hold H*W rows.  x = (w - cx) / fx * z in fp32, the reference's operation order.
 * ---------------------------------------------------------------------------------------------- */
size_t endo_point_cloud_workspace_bytes(int H, int W);
int endo_point_cloud_from_depth(const float* depth, const unsigned char* color_bgr, const float* mask, float fx, float fy,
                                float cx, float cy, int H, int W, int downsampling, int use_threshold, float min_threshold,
                                float max_threshold, float* points, int* count, void* ws, size_t ws_bytes,
                                endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * utils.get_torch_training_data (utils.py:460-612): sparse SfM rasteriser of one image pair.  points[M][4] homogeneous
 * (double), projections[2][3][4], extrinsics[2][4][4] (double), visibility[2][M] (the two columns of view_indexes_per_point,
 * > 0.5 = visible), clean[M] or NULL, mask_boundary[H][W] uint8 (255 = inside) -> the four pairs of images the reference
 * returns: depth_mask[2][H][W], depth[2][H][W], flow_mask[2][H][W], flow[2][H][W][2].  float64 projection, np.round,
 * last-point-wins scatter, float32 flow normalisation and the |flow| > 5 outlier rule exactly as in the reference.
 * ---------------------------------------------------------------------------------------------- */
size_t endo_rasterize_workspace_bytes(int M, int H, int W);
int endo_rasterize_pair(const double* points, const double* projections, const double* extrinsics, const float* visibility,
                        const float* clean, const unsigned char* mask_boundary, int M, int H, int W, float* depth_mask,
                        float* depth, float* flow_mask, float* flow, void* ws, size_t ws_bytes, endo_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * utils.get_pair_color_imgs after the JPEG decode (utils.py:446-452) + dataset.py:148,446-453's normalisation, one image:
 * src_bgr[src_h][src_w][3] uint8 (what cv2.imread returns) -> cv2.resize(fx = fy = 1/downsampling, INTER_LINEAR) bit for bit
 * -> crop [start_h:end_h, start_w:end_w] -> optional BGR->RGB -> out_u8[H][W][3] and / or out_norm[3][H][W] float32
 * ((v - 127.5) * (1/127.5): albumentations Normalize(0.5, 0.5, 255) + img_to_tensor).  Either output may be NULL.
 * ENDO_ERR_BAD_SHAPE: crop outside the resized image; ENDO_ERR_CONFIG: factor 2 on an odd-sized image (cv::resize takes a
 * different border path there).
 * ---------------------------------------------------------------------------------------------- */
int endo_resize_crop_u8(const unsigned char* src_bgr, int src_h, int src_w, double downsampling, int start_h, int end_h,
                        int start_w, int end_w, int swap_rb, unsigned char* out_u8, float* out_norm, endo_stream_t stream);

/* TMA bring-up probe (tests only): out[box_h][box_w][box_c] = the box of the NHWC fp32 buffer src[B][H][W][C] at channel c0,
 * column x0, row y0 (either may be negative or overhang: zero fill) of image b, loaded by one cp.async.bulk.tensor. */
int endo_tma_probe(const float* src, int B, int H, int W, int C, int box_c, int box_w, int box_h, int c0, int x0,
                   int y0, int b, float* out, endo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ENDO_B200_H */
