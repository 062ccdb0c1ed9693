// Training losses of the reference as single-launch reductions for sm_100a:
//   SparseMaskedL1Loss      (/root/reference/losses.py:57-66)
//   NormalizedDistanceLoss  (/root/reference/losses.py:112-146)
//   ScaleInvariantLoss      (/root/reference/losses.py:17-32)
// Forward: one pass over the maps (128-bit loads), per-block fp64 partials, the last block to arrive
// reduces them in a fixed order and writes the scalar loss plus the per-sample sums the backward needs.
// Backward: one elementwise pass; the upstream gradient is read from device memory (no host sync).
#include "common.cuh"

namespace endo {

constexpr int kLT = 256;

__host__ __device__ inline int loss_nblk(int HW) {
    int n = (HW + kLT * 4 - 1) / (kLT * 4);
    return n < 64 ? n : 64;
}

template <int VEC>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        float4 q = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        v[0] = __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else p[0] = v[0];
}
__device__ __forceinline__ float sgn(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

template <int N>
__device__ __forceinline__ void write_partials(double* partials, int b, int nblk, const double (&v)[N]) {
    if (threadIdx.x == 0)
        for (int i = 0; i < N; ++i) partials[((size_t)b * nblk + blockIdx.x) * N + i] = v[i];
}
template <int N>
__device__ __forceinline__ void sum_partials(const double* partials, int b, int nblk, double (&v)[N]) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    for (int k = 0; k < nblk; ++k)
        for (int i = 0; i < N; ++i) v[i] += ld_cg(partials + ((size_t)b * nblk + k) * N + i);
}
// batch mean of per-sample losses computed by threads bb = threadIdx.x, fixed order
__device__ __forceinline__ void finish_mean(double local, int B, float* loss) {
    __shared__ double s_loss[kLT];
    s_loss[threadIdx.x] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int k = 0; k < kLT; ++k) a += s_loss[k];
        loss[0] = (float)(a / B);
    }
}

// ------------------------------------------------------------------------------------------ L1
template <int VEC>
__global__ void __launch_bounds__(kLT)
sparse_l1_fwd_kernel(const float* __restrict__ f, const float* __restrict__ fd, const float* __restrict__ m,
                     unsigned* counter, double* __restrict__ partials, float* __restrict__ loss,
                     float* __restrict__ stats, int B, int HW, float eps) {
    __shared__ double red[2 * kLT / 32];
    const int b = blockIdx.y, nblk = gridDim.x;
    float s0 = 0.f, s1 = 0.f;
    for (int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC; p0 < HW; p0 += nblk * kLT * VEC) {
        float a0[VEC], a1[VEC], c0[VEC], c1[VEC], mm[VEC];
        ldv<VEC>(f + ((size_t)b * 2 + 0) * HW + p0, a0);
        ldv<VEC>(f + ((size_t)b * 2 + 1) * HW + p0, a1);
        ldv<VEC>(fd + ((size_t)b * 2 + 0) * HW + p0, c0);
        ldv<VEC>(fd + ((size_t)b * 2 + 1) * HW + p0, c1);
        ldv<VEC>(m + (size_t)b * HW + p0, mm);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            s0 += mm[i] * fabsf(a0[i] - c0[i]) + mm[i] * fabsf(a1[i] - c1[i]);   // losses.py:64
            s1 += mm[i];
        }
    }
    double v[2] = {(double)s0, (double)s1};
    block_sum<2, kLT>(v, red);
    write_partials<2>(partials, b, nblk, v);
    if (arrive_is_last(counter, gridDim.x * gridDim.y)) {
        double local = 0.0;
        for (int bb = threadIdx.x; bb < B; bb += kLT) {
            double t[2];
            sum_partials<2>(partials, bb, nblk, t);
            stats[bb * 2 + 0] = (float)t[0]; stats[bb * 2 + 1] = (float)t[1];
            local += t[0] / ((double)eps + t[1]);                                   // :64-65
        }
        finish_mean(local, B, loss);                                                // :66
    }
}

template <int VEC>
__global__ void __launch_bounds__(kLT)
sparse_l1_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ f, const float* __restrict__ fd,
                     const float* __restrict__ m, const float* __restrict__ stats, float* __restrict__ g_fd,
                     float* __restrict__ g_f, int B, int HW, float eps) {
    const int b = blockIdx.y;
    const int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    const float coef = g_loss[0] / (float)B / (eps + stats[b * 2 + 1]);
    float mm[VEC];
    ldv<VEC>(m + (size_t)b * HW + p0, mm);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        float a[VEC], q[VEC], o[VEC], o2[VEC];
        ldv<VEC>(f + ((size_t)b * 2 + c) * HW + p0, a);
        ldv<VEC>(fd + ((size_t)b * 2 + c) * HW + p0, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            o2[i] = coef * mm[i] * sgn(a[i] - q[i]);
            o[i] = -o2[i];
        }
        stv<VEC>(g_fd + ((size_t)b * 2 + c) * HW + p0, o);
        if (g_f) stv<VEC>(g_f + ((size_t)b * 2 + c) * HW + p0, o2);
    }
}

// ------------------------------------------------------------------------------------------ NDL
template <int VEC>
__global__ void __launch_bounds__(kLT)
norm_dist_fwd_kernel(const float* __restrict__ d, const float* __restrict__ w, const float* __restrict__ m,
                     const float* __restrict__ K, unsigned* counter, double* __restrict__ partials,
                     float* __restrict__ loss, float* __restrict__ stats, int B, int H, int W, float eps) {
    __shared__ double red[4 * kLT / 32];
    const int b = blockIdx.y, nblk = gridDim.x, HW = H * W;
    const float fx = K[b * 9 + 0], fy = K[b * 9 + 4], cx = K[b * 9 + 2], cy = K[b * 9 + 5];   // losses.py:124-127
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC; p0 < HW; p0 += nblk * kLT * VEC) {
        float dd[VEC], ww[VEC], mm[VEC];
        ldv<VEC>(d + (size_t)b * HW + p0, dd);
        ldv<VEC>(w + (size_t)b * HW + p0, ww);
        ldv<VEC>(m + (size_t)b * HW + p0, mm);
        const int y = p0 / W, x0 = p0 - y * W;
        const float ay = ((float)y - cy) / fy;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float ax = ((float)(x0 + i) - cx) / fx;          // VEC == 4 only when W % 4 == 0: one row
            const float e = fabsf(ax * dd[i] - ax * ww[i]) + fabsf(ay * dd[i] - ay * ww[i]) + fabsf(dd[i] - ww[i]);
            s0 += mm[i] * e;                                        // :141
            s1 += mm[i] * (dd[i] + fabsf(ww[i]));                   // :144
            s2 += mm[i] * dd[i];                                    // :130
            s3 += mm[i];
        }
    }
    double v[4] = {(double)s0, (double)s1, (double)s2, (double)s3};
    block_sum<4, kLT>(v, red);
    write_partials<4>(partials, b, nblk, v);
    if (arrive_is_last(counter, gridDim.x * gridDim.y)) {
        double local = 0.0;
        for (int bb = threadIdx.x; bb < B; bb += kLT) {
            double t[4];
            sum_partials<4>(partials, bb, nblk, t);
            const double mean = t[2] / ((double)eps + t[3]);        // :130-132 (no_grad)
            const double den = 1.0e-5 * mean + t[1];                // :143
            stats[bb * 4 + 0] = (float)t[0]; stats[bb * 4 + 1] = (float)den;
            stats[bb * 4 + 2] = (float)mean; stats[bb * 4 + 3] = (float)t[3];
            local += 2.0 * t[0] / den;
        }
        finish_mean(local, B, loss);                                // :146
    }
}

template <int VEC>
__global__ void __launch_bounds__(kLT)
norm_dist_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ d, const float* __restrict__ w,
                     const float* __restrict__ m, const float* __restrict__ K, const float* __restrict__ stats,
                     float* __restrict__ g_d, float* __restrict__ g_w, int B, int H, int W) {
    const int b = blockIdx.y, HW = H * W;
    const int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    const float fx = K[b * 9 + 0], fy = K[b * 9 + 4], cx = K[b * 9 + 2], cy = K[b * 9 + 5];
    const float num = stats[b * 4 + 0], den = stats[b * 4 + 1];
    const float c = g_loss[0] / (float)B;
    const float k1 = c * 2.0f / den;                                // d L / d numerator
    const float k2 = c * 2.0f * num / (den * den);                  // -d L / d denominator
    float dd[VEC], ww[VEC], mm[VEC], od[VEC], ow[VEC];
    ldv<VEC>(d + (size_t)b * HW + p0, dd);
    ldv<VEC>(w + (size_t)b * HW + p0, ww);
    ldv<VEC>(m + (size_t)b * HW + p0, mm);
    const int y = p0 / W, x0 = p0 - y * W;
    const float ay = ((float)y - cy) / fy;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float ax = ((float)(x0 + i) - cx) / fx;
        const float e = sgn(ax * dd[i] - ax * ww[i]) * ax + sgn(ay * dd[i] - ay * ww[i]) * ay + sgn(dd[i] - ww[i]);
        od[i] = mm[i] * (k1 * e - k2);
        ow[i] = mm[i] * (-k1 * e - k2 * sgn(ww[i]));
    }
    stv<VEC>(g_d + (size_t)b * HW + p0, od);
    stv<VEC>(g_w + (size_t)b * HW + p0, ow);
}

// ------------------------------------------------------------------------------------------ SIL
template <int VEC>
__global__ void __launch_bounds__(kLT)
scale_inv_fwd_kernel(const float* __restrict__ p, const float* __restrict__ g, const float* __restrict__ bnd,
                     unsigned* counter, double* __restrict__ partials, float* __restrict__ loss,
                     float* __restrict__ stats, int B, int HW, float eps) {
    __shared__ double red[3 * kLT / 32];
    const int b = blockIdx.y, nblk = gridDim.x;
    double v[3] = {0.0, 0.0, 0.0};
    for (int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC; p0 < HW; p0 += nblk * kLT * VEC) {
        float pp[VEC], gg[VEC], bb[VEC];
        ldv<VEC>(p + (size_t)b * HW + p0, pp);
        ldv<VEC>(g + (size_t)b * HW + p0, gg);
        ldv<VEC>(bnd + (size_t)b * HW + p0, bb);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float r = logf(bb[i] * pp[i] + eps) - logf(bb[i] * gg[i] + eps);   // losses.py:24-25
            s0 += r * r; s1 += r; s2 += bb[i];
        }
        v[0] += (double)s0; v[1] += (double)s1; v[2] += (double)s2;
    }
    block_sum<3, kLT>(v, red);
    write_partials<3>(partials, b, nblk, v);
    if (arrive_is_last(counter, gridDim.x * gridDim.y)) {
        double local = 0.0;
        for (int bb = threadIdx.x; bb < B; bb += kLT) {
            double t[3];
            sum_partials<3>(partials, bb, nblk, t);
            stats[bb * 4 + 0] = (float)t[0]; stats[bb * 4 + 1] = (float)t[1];
            stats[bb * 4 + 2] = (float)t[2]; stats[bb * 4 + 3] = 0.0f;
            local += t[0] / t[2] + (t[1] * t[1]) / (t[2] * t[2]);                      // :27-31
        }
        finish_mean(local, B, loss);                                                   // :32
    }
}

template <int VEC>
__global__ void __launch_bounds__(kLT)
scale_inv_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ p, const float* __restrict__ g,
                     const float* __restrict__ bnd, const float* __restrict__ stats, float* __restrict__ g_p,
                     float* __restrict__ g_g, int B, int HW, float eps) {
    const int b = blockIdx.y;
    const int p0 = (blockIdx.x * kLT + threadIdx.x) * VEC;
    if (p0 >= HW) return;
    const float wsum = stats[b * 4 + 2], sr = stats[b * 4 + 1];
    const float c = g_loss[0] / (float)B;
    const float k1 = c * 2.0f / wsum, k2 = c * 2.0f * sr / (wsum * wsum);
    float pp[VEC], gg[VEC], bb[VEC], o1[VEC], o2[VEC];
    ldv<VEC>(p + (size_t)b * HW + p0, pp);
    ldv<VEC>(g + (size_t)b * HW + p0, gg);
    ldv<VEC>(bnd + (size_t)b * HW + p0, bb);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float ap = bb[i] * pp[i] + eps, ag = bb[i] * gg[i] + eps;
        const float r = logf(ap) - logf(ag);
        const float dr = k1 * r + k2;                               // d loss / d r
        o1[i] = dr * bb[i] / ap;
        o2[i] = -dr * bb[i] / ag;
    }
    stv<VEC>(g_p + (size_t)b * HW + p0, o1);
    if (g_g) stv<VEC>(g_g + (size_t)b * HW + p0, o2);
}

}  // namespace endo

using namespace endo;

static inline bool v4ok(int HW, int W, std::initializer_list<const void*> ptrs) {
    if ((HW & 3) || (W & 3)) return false;
    for (const void* p : ptrs)
        if (p && !aligned16(p)) return false;
    return true;
}
#define REQ_DIMS(B, H, W) \
    if ((B) <= 0 || (H) <= 0 || (W) <= 0 || (long long)(H) * (W) > (1ll << 30) || (B) > 65535) return ENDO_ERR_BAD_SHAPE
#define REQ(p) \
    if ((p) == nullptr) return ENDO_ERR_BAD_POINTER

extern "C" size_t endo_loss_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return ENDO_WS_HEADER_BYTES + sizeof(double) * (size_t)B * loss_nblk(H * W) * 4 + 64;
}

#define LOSS_WS()                                                                                     \
    if (!ws || ws_bytes < endo_loss_workspace_bytes(B, H, W) || !aligned16(ws)) return ENDO_ERR_WORKSPACE; \
    unsigned* counter = reinterpret_cast<unsigned*>(ws);                                              \
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + ENDO_WS_HEADER_BYTES); \
    cudaStream_t s = (cudaStream_t)stream;                                                            \
    ProfScope prof(PC_LOSS, s);                                                                       \
    const int HW = H * W;                                                                             \
    dim3 rgrid(loss_nblk(HW), B)

extern "C" int endo_sparse_l1_fwd(const float* flows, const float* flows_from_depth, const float* masks, float* loss,
                                  float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                                  endo_stream_t stream) {
    REQ_DIMS(B, H, W); REQ(flows); REQ(flows_from_depth); REQ(masks); REQ(loss); REQ(stats);
    LOSS_WS();
    if (v4ok(HW, W, {flows, flows_from_depth, masks}))
        sparse_l1_fwd_kernel<4><<<rgrid, kLT, 0, s>>>(flows, flows_from_depth, masks, counter, partials, loss, stats, B, HW, eps);
    else
        sparse_l1_fwd_kernel<1><<<rgrid, kLT, 0, s>>>(flows, flows_from_depth, masks, counter, partials, loss, stats, B, HW, eps);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_sparse_l1_bwd(const float* g_loss, const float* flows, const float* flows_from_depth,
                                  const float* masks, const float* stats, float* g_flows_from_depth, float* g_flows,
                                  int B, int H, int W, float eps, endo_stream_t stream) {
    REQ_DIMS(B, H, W); REQ(g_loss); REQ(flows); REQ(flows_from_depth); REQ(masks); REQ(stats); REQ(g_flows_from_depth);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_LOSS, s);
    const int HW = H * W;
    if (v4ok(HW, W, {flows, flows_from_depth, masks, g_flows_from_depth, g_flows}))
        sparse_l1_bwd_kernel<4><<<dim3(cdiv(HW, kLT * 4), B), kLT, 0, s>>>(g_loss, flows, flows_from_depth, masks, stats,
                                                                         g_flows_from_depth, g_flows, B, HW, eps);
    else
        sparse_l1_bwd_kernel<1><<<dim3(cdiv(HW, kLT), B), kLT, 0, s>>>(g_loss, flows, flows_from_depth, masks, stats,
                                                                     g_flows_from_depth, g_flows, B, HW, eps);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_norm_dist_fwd(const float* depth, const float* warped, const float* intersect, const float* K,
                                  float* loss, float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                                  endo_stream_t stream) {
    REQ_DIMS(B, H, W); REQ(depth); REQ(warped); REQ(intersect); REQ(K); REQ(loss); REQ(stats);
    LOSS_WS();
    if (v4ok(HW, W, {depth, warped, intersect}))
        norm_dist_fwd_kernel<4><<<rgrid, kLT, 0, s>>>(depth, warped, intersect, K, counter, partials, loss, stats, B, H, W, eps);
    else
        norm_dist_fwd_kernel<1><<<rgrid, kLT, 0, s>>>(depth, warped, intersect, K, counter, partials, loss, stats, B, H, W, eps);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_norm_dist_bwd(const float* g_loss, const float* depth, const float* warped, const float* intersect,
                                  const float* K, const float* stats, float* g_depth, float* g_warped, int B, int H,
                                  int W, float eps, endo_stream_t stream) {
    (void)eps;
    REQ_DIMS(B, H, W); REQ(g_loss); REQ(depth); REQ(warped); REQ(intersect); REQ(K); REQ(stats); REQ(g_depth); REQ(g_warped);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_LOSS, s);
    const int HW = H * W;
    if (v4ok(HW, W, {depth, warped, intersect, g_depth, g_warped}))
        norm_dist_bwd_kernel<4><<<dim3(cdiv(HW, kLT * 4), B), kLT, 0, s>>>(g_loss, depth, warped, intersect, K, stats,
                                                                         g_depth, g_warped, B, H, W);
    else
        norm_dist_bwd_kernel<1><<<dim3(cdiv(HW, kLT), B), kLT, 0, s>>>(g_loss, depth, warped, intersect, K, stats,
                                                                     g_depth, g_warped, B, H, W);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_scale_inv_fwd(const float* pred, const float* goal, const float* boundaries, float* loss,
                                  float* stats, int B, int H, int W, float eps, void* ws, size_t ws_bytes,
                                  endo_stream_t stream) {
    REQ_DIMS(B, H, W); REQ(pred); REQ(goal); REQ(boundaries); REQ(loss); REQ(stats);
    LOSS_WS();
    if (v4ok(HW, W, {pred, goal, boundaries}))
        scale_inv_fwd_kernel<4><<<rgrid, kLT, 0, s>>>(pred, goal, boundaries, counter, partials, loss, stats, B, HW, eps);
    else
        scale_inv_fwd_kernel<1><<<rgrid, kLT, 0, s>>>(pred, goal, boundaries, counter, partials, loss, stats, B, HW, eps);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_scale_inv_bwd(const float* g_loss, const float* pred, const float* goal, const float* boundaries,
                                  const float* stats, float* g_pred, float* g_goal, int B, int H, int W, float eps,
                                  endo_stream_t stream) {
    REQ_DIMS(B, H, W); REQ(g_loss); REQ(pred); REQ(goal); REQ(boundaries); REQ(stats); REQ(g_pred);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_LOSS, s);
    const int HW = H * W;
    if (v4ok(HW, W, {pred, goal, boundaries, g_pred, g_goal}))
        scale_inv_bwd_kernel<4><<<dim3(cdiv(HW, kLT * 4), B), kLT, 0, s>>>(g_loss, pred, goal, boundaries, stats, g_pred,
                                                                         g_goal, B, HW, eps);
    else
        scale_inv_bwd_kernel<1><<<dim3(cdiv(HW, kLT), B), kLT, 0, s>>>(g_loss, pred, goal, boundaries, stats, g_pred,
                                                                     g_goal, B, HW, eps);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}
