// Optimiser tail of the reference step (/root/reference/train.py:327-328):
//   torch.nn.utils.clip_grad_norm_(params, 10.0)  +  torch.optim.SGD(momentum=0.9).step()
// on ONE flat parameter / gradient / momentum array (1,374,865 floats for FCDenseNet57) instead of
// foreach kernels over 210 tensors.  Two launches: fp64 block partials of sum(g^2) with a fixed-order
// final reduction (deterministic), then the fused clip + momentum + update pass.
#include "common.cuh"

namespace endo {

constexpr int kOT = 256;

__global__ void __launch_bounds__(kOT)
grad_sqnorm_kernel(const float* __restrict__ g, long long n, unsigned* counter, double* __restrict__ partials,
                   double* __restrict__ total, float* __restrict__ norm_out) {
    __shared__ double red[kOT / 32];
    double v[1] = {0.0};
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float acc = 0.f;
    int cnt = 0;
    for (long long i = (long long)blockIdx.x * kOT + threadIdx.x; i < n4; i += (long long)gridDim.x * kOT) {
        const float4 q = __ldg(g4 + i);
        acc += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        if (++cnt == 16) { v[0] += (double)acc; acc = 0.f; cnt = 0; }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * kOT + threadIdx.x; i < n; i += (long long)gridDim.x * kOT) {
        const float q = g[i];
        acc += q * q;
    }
    v[0] += (double)acc;
    block_sum<1, kOT>(v, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = v[0];
    if (arrive_is_last(counter, gridDim.x)) {
        if (threadIdx.x == 0) {
            double a = 0.0;
            for (unsigned k = 0; k < gridDim.x; ++k) a += ld_cg(partials + k);
            total[0] = a;
            if (norm_out) norm_out[0] = (float)sqrt(a);
        }
    }
}

__global__ void __launch_bounds__(kOT)
sgd_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf, long long n, float lr,
                float momentum, float max_norm, int first_step, const float* __restrict__ finite_flag,
                const double* __restrict__ total, const float* __restrict__ lr_dev) {
    if (finite_flag && finite_flag[0] == 0.0f) return;            // device-side NaN guard (train.py:317-322)
    if (lr_dev) lr = lr_dev[0];                                   // learning rate as a device scalar (CUDA-graph replay + LR schedule)
    const float norm = (float)sqrt(total[0]);
    const float raw = max_norm / (norm + 1.0e-6f);                // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    const float coef = (raw > 1.0f) ? 1.0f : raw;                 // a NaN norm stays NaN (torch.clamp keeps NaN)
    const long long i0 = ((long long)blockIdx.x * kOT + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (i0 + 3 < n) {
        float4 gq = *reinterpret_cast<float4*>(g + i0);
        float4 pq = *reinterpret_cast<float4*>(p + i0);
        float4 bq = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(buf + i0);
        gq.x *= coef; gq.y *= coef; gq.z *= coef; gq.w *= coef;
        if (first_step) bq = gq;
        else { bq.x = momentum * bq.x + gq.x; bq.y = momentum * bq.y + gq.y; bq.z = momentum * bq.z + gq.z; bq.w = momentum * bq.w + gq.w; }
        pq.x -= lr * bq.x; pq.y -= lr * bq.y; pq.z -= lr * bq.z; pq.w -= lr * bq.w;
        *reinterpret_cast<float4*>(g + i0) = gq;                  // clip_grad_norm_ scales .grad in place
        *reinterpret_cast<float4*>(buf + i0) = bq;
        *reinterpret_cast<float4*>(p + i0) = pq;
    } else {
        for (long long i = i0; i < n; ++i) {
            const float gq = g[i] * coef;
            const float bq = first_step ? gq : momentum * buf[i] + gq;
            g[i] = gq; buf[i] = bq; p[i] -= lr * bq;
        }
    }
}

}  // namespace endo

using namespace endo;

static inline int sgd_blocks(long long n) {
    long long b = (n / 4 + kOT - 1) / kOT;
    if (b < 1) b = 1;
    return (int)(b < 4 * kNumSMs ? b : 4 * kNumSMs);
}

extern "C" size_t endo_sgd_workspace_bytes(long long n) {
    if (n <= 0) return 0;
    return ENDO_WS_HEADER_BYTES + sizeof(double) * (size_t)(sgd_blocks(n) + 2) + 64;
}

static int sgd_clip_step(float* params, float* grads, float* momentum_buf, long long n, float lr, const float* lr_dev,
                         float momentum, float max_norm, int first_step, const float* finite_flag,
                         float* grad_norm_out, void* ws, size_t ws_bytes, endo_stream_t stream) {
    if (n <= 0) return ENDO_ERR_BAD_SHAPE;
    if (!params || !grads || !momentum_buf) return ENDO_ERR_BAD_POINTER;
    if (!aligned16(params) || !aligned16(grads) || !aligned16(momentum_buf)) return ENDO_ERR_BAD_POINTER;
    if (!ws || ws_bytes < endo_sgd_workspace_bytes(n) || !aligned16(ws)) return ENDO_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(PC_OPT, s);
    unsigned* counter = reinterpret_cast<unsigned*>(ws) + 8;    // separate ticket from the loss kernels
    double* total = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + ENDO_WS_HEADER_BYTES);
    double* partials = total + 2;
    const int nb = sgd_blocks(n);
    grad_sqnorm_kernel<<<nb, kOT, 0, s>>>(grads, n, counter, partials, total, grad_norm_out);
    ENDO_CHECK_LAUNCH();
    const int nb2 = cdiv(cdiv(n, 4), kOT);
    sgd_clip_kernel<<<nb2, kOT, 0, s>>>(params, grads, momentum_buf, n, lr, momentum, max_norm, first_step,
                                        finite_flag, total, lr_dev);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

extern "C" int endo_sgd_clip_step(float* params, float* grads, float* momentum_buf, long long n, float lr,
                                  float momentum, float max_norm, int first_step, const float* finite_flag,
                                  float* grad_norm_out, void* ws, size_t ws_bytes, endo_stream_t stream) {
    return sgd_clip_step(params, grads, momentum_buf, n, lr, nullptr, momentum, max_norm, first_step, finite_flag,
                         grad_norm_out, ws, ws_bytes, stream);
}

extern "C" int endo_sgd_clip_step_dev(float* params, float* grads, float* momentum_buf, long long n, const float* lr_dev,
                                      float momentum, float max_norm, const float* finite_flag, float* grad_norm_out,
                                      void* ws, size_t ws_bytes, endo_stream_t stream) {
    if (!lr_dev) return ENDO_ERR_BAD_POINTER;
    return sgd_clip_step(params, grads, momentum_buf, n, 0.f, lr_dev, momentum, max_norm, 0, finite_flag, grad_norm_out, ws,
                         ws_bytes, stream);
}
