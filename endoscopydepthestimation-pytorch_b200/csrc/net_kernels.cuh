// Device kernels of the FC-DenseNet engine, fp32 FFMA path (ENDO_MATH_FP32).
//
// One templated implicit-GEMM convolution kernel serves every convolution of the forward pass and every
// data-gradient of the backward pass; what changes is
//   * the LOADER that produces the GEMM A operand on the fly while staging it into shared memory
//       LM_BNRELU   relu(a_c * x + b_c)          BatchNorm(train)+ReLU fused into the operand load
//       LM_PLAIN    x (optionally nearest-x2 upsampled: TransitionUp, models.py:73)
//       LM_NCHW     x read from the user's NCHW input tensor (firstconv)
//       LM_GRAD     g + A_c + B_c * x            output gradient with the lazy BN-backward correction
//       LM_GRADPOOL the same routed through MaxPool2d's argmax (TransitionDown backward)
//   * the EPILOGUE
//       EM_STORE      + bias, store 12/16/48 channels in place, emit per-channel sum / sum-of-squares
//       EM_POOL       + bias, 2x2 max-pool (first-max-wins like ATen), store, argmax byte, statistics
//       EM_DGRAD_BN   ReLU mask, BN-backward sums, scaled accumulate into the gradient buffer
//       EM_DGRAD_UP   2x2 sum (backward of nearest upsampling), accumulate
//       EM_PARTIAL    split-K: blockIdx.y owns a slice of the input channels and stores its raw partial sums to a
//                     scratch tensor; splitk_finish_kernel adds the slices in a fixed order (low-resolution levels,
//                     where a CTA per 32x32 tile over ALL input channels would leave most SMs idle)
//   * the weight view: forward (k = ci, n = co) or data-gradient (k = co, n = ci, taps flipped).
// Thread mapping: a warp owns PX output rows x 32 consecutive columns (lane = column, so every shared
// memory access is either conflict-free or a broadcast), a CTA stacks NW warps vertically; each
// thread keeps PX x CO fp32 accumulators.
#pragma once
#include "common.cuh"

namespace endo {

enum { LM_PLAIN = 0, LM_BNRELU = 1, LM_NCHW = 2, LM_GRAD = 3, LM_GRADPOOL = 4 };
enum { EM_STORE = 0, EM_POOL = 1, EM_DGRAD_BN = 2, EM_DGRAD_UP = 3, EM_PARTIAL = 4 };
enum { WM_FWD = 0, WM_DGRAD = 1 };

constexpr int KC = 8;   // input channels staged per main-loop step

struct ConvArgs {
    // ---- operand A (loader)
    const float* in;          // activation / gradient buffer the loader reads (NHWC, or NCHW for LM_NCHW)
    const float* in2;         // LM_GRAD*: activation buffer x matching `in` (lazy correction needs x)
    const float* in_ab;       // LM_GRAD*: [G][in_C][2] (A, Bc) of that buffer
    const float* coef;        // LM_BNRELU: [G][K][4] (a = gamma*invstd, beta, mean, invstd) of this layer's BatchNorm
    const unsigned char* argmax;  // LM_GRADPOOL: [B, ih, iw, K] window index chosen by the forward pool
    int in_C, in_off, K;      // channel stride of the buffer, first channel, number of GEMM-K channels
    int ih, iw;               // spatial size of the buffer the loader reads
    // ---- weights
    const float* w;           // OIHW tensor of the layer
    const float* bias;
    int w_cin;                // I of OIHW
    // ---- output
    float* out;               // buffer written / accumulated (NHWC)
    int out_C, out_off, N;    // channel stride, first channel, number of output channels
    int oh, ow;               // output spatial size == GEMM pixel grid
    int B, G;
    double* stats;            // EM_STORE/EM_POOL: [G][out_C][2]; EM_DGRAD_BN: [G][maxC][2] (index = n)
    int stats_C;
    unsigned char* argmax_out;    // EM_POOL: [B, oh/2, ow/2, N]
    // ---- EM_DGRAD_BN
    const float* x;           // activation buffer (same geometry as `out`)
    const float* ep_coef;     // [G][N][4] (a, beta, mean, invstd) of the BatchNorm being differentiated
    // ---- EM_PARTIAL
    float* partial;           // [ksplit][B][oh][ow][CO]
    int ksplit;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------------------
template <int LM, bool UP>
__device__ __forceinline__ float4 load_a4(const ConvArgs& A, int b, int g, int y, int x, int c) {
    // (y, x) are coordinates in the OUTPUT pixel grid; returns channels c..c+3 of operand A (zeros outside)
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < 0 || y >= A.oh || x < 0 || x >= A.ow || c >= A.K) return r;
    if constexpr (LM == LM_PLAIN) {
        const int sy = UP ? (y >> 1) : y, sx = UP ? (x >> 1) : x;
        return ldg4(A.in + ((size_t)(b * A.ih + sy) * A.iw + sx) * A.in_C + A.in_off + c);
    } else if constexpr (LM == LM_BNRELU) {
        // relu(a * (x - mean) + beta): subtracting the mean first keeps the pre-activation as well conditioned as the
        // reference's (x - mean) / sqrt(var + eps) * gamma + beta, so ReLU masks agree with it up to fp32 rounding
        const float4 v = ldg4(A.in + ((size_t)(b * A.ih + y) * A.iw + x) * A.in_C + A.in_off + c);
        const float* cf = A.coef + ((size_t)g * A.K + c) * 4;
        const float4 c0 = ldg4(cf), c1 = ldg4(cf + 4), c2 = ldg4(cf + 8), c3 = ldg4(cf + 12);
        r.x = fmaxf(fmaf(c0.x, v.x - c0.z, c0.y), 0.f); r.y = fmaxf(fmaf(c1.x, v.y - c1.z, c1.y), 0.f);
        r.z = fmaxf(fmaf(c2.x, v.z - c2.z, c2.y), 0.f); r.w = fmaxf(fmaf(c3.x, v.w - c3.z, c3.y), 0.f);
        return r;
    } else if constexpr (LM == LM_GRAD) {
        const size_t o = ((size_t)(b * A.ih + y) * A.iw + x) * A.in_C + A.in_off + c;
        const float4 gq = ldg4(A.in + o), xq = ldg4(A.in2 + o);
        const float* ab = A.in_ab + ((size_t)g * A.in_C + A.in_off + c) * 2;
        const float4 c0 = ldg4(ab), c1 = ldg4(ab + 4);
        r.x = gq.x + fmaf(c0.y, xq.x, c0.x); r.y = gq.y + fmaf(c0.w, xq.y, c0.z);
        r.z = gq.z + fmaf(c1.y, xq.z, c1.x); r.w = gq.w + fmaf(c1.w, xq.w, c1.z);
        return r;
    } else if constexpr (LM == LM_GRADPOOL) {
        const int sy = y >> 1, sx = x >> 1;
        const unsigned pos = ((y & 1) << 1) | (x & 1);
        const size_t pp = (size_t)(b * A.ih + sy) * A.iw + sx;
        const unsigned am = __ldg(reinterpret_cast<const unsigned*>(A.argmax + pp * A.K + c));
        const bool h0 = (am & 0xffu) == pos, h1 = ((am >> 8) & 0xffu) == pos, h2 = ((am >> 16) & 0xffu) == pos,
                   h3 = (am >> 24) == pos;
        if (!(h0 | h1 | h2 | h3)) return r;
        const size_t o = pp * A.in_C + A.in_off + c;
        const float4 gq = ldg4(A.in + o), xq = ldg4(A.in2 + o);
        const float* ab = A.in_ab + ((size_t)g * A.in_C + A.in_off + c) * 2;
        const float4 c0 = ldg4(ab), c1 = ldg4(ab + 4);
        if (h0) r.x = gq.x + fmaf(c0.y, xq.x, c0.x);
        if (h1) r.y = gq.y + fmaf(c0.w, xq.y, c0.z);
        if (h2) r.z = gq.z + fmaf(c1.y, xq.z, c1.x);
        if (h3) r.w = gq.w + fmaf(c1.w, xq.w, c1.z);
        return r;
    }
    return r;
}

template <int KS, int PX, int NW>
struct ConvTile {
    static constexpr int PAD = KS / 2;
    static constexpr int TR = NW * PX;             // output rows per CTA
    static constexpr int TRP = TR + 2 * PAD;
    static constexpr int TWP = 32 + 2 * PAD;
    static constexpr int PLANE = TRP * TWP + ((TRP * TWP) % 32 == 4 ? 0 : ((36 - (TRP * TWP) % 32) % 32));   // plane % 32 == 4
};

template <int KS, int PX, int CO, int NW, int LM, int EM, int WM, bool UP, int KCT = KC, int MINB = 1, bool PF = false>
__global__ void __launch_bounds__(NW * 32, MINB)
conv_kernel(const ConvArgs A) {
    pdl_enter();
    using T = ConvTile<KS, PX, NW>;
    constexpr int PAD = T::PAD, TR = T::TR, TRP = T::TRP, TWP = T::TWP, PLANE = T::PLANE, TAPS = KS * KS;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) float smem[];
    float* a_s = smem;                          // [KCT][PLANE]
    float* w_s = smem + KCT * PLANE;             // [KCT][TAPS][CO]

    const int tiles_x = (A.ow + 31) / 32;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * TR, x0 = tx * 32;
    const int n0 = (EM == EM_PARTIAL) ? 0 : blockIdx.y * CO;
    const int b = blockIdx.z;
    const int g = b / (A.B / A.G);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = warp * PX;

    float acc[PX][CO];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;

    int k_lo = 0, k_hi = A.K;
    if constexpr (EM == EM_PARTIAL) {
        const int kper = ((A.K + A.ksplit - 1) / A.ksplit + KCT - 1) / KCT * KCT;
        k_lo = blockIdx.y * kper;
        k_hi = (k_lo + kper < A.K) ? (k_lo + kper) : A.K;
    }
    auto compute_chunk = [&]() {
        // ---- main loop: PX x CO register tile per thread
#pragma unroll 1
        for (int kk = 0; kk < KCT; ++kk) {
            const float* ap = a_s + kk * PLANE + r0 * TWP + lane;
            const float* wp = w_s + kk * TAPS * CO;
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                float av[PX + KS - 1];
#pragma unroll
                for (int i = 0; i < PX + KS - 1; ++i) av[i] = ap[i * TWP + kx];
#pragma unroll
                for (int ky = 0; ky < KS; ++ky) {
                    float wv[CO];
#pragma unroll
                    for (int j = 0; j < CO / 4; ++j) {
                        const float4 q = *reinterpret_cast<const float4*>(wp + (ky * KS + kx) * CO + j * 4);
                        wv[j * 4] = q.x; wv[j * 4 + 1] = q.y; wv[j * 4 + 2] = q.z; wv[j * 4 + 3] = q.w;
                    }
#pragma unroll
                    for (int i = 0; i < PX; ++i)
#pragma unroll
                        for (int j = 0; j < CO; ++j) acc[i][j] = fmaf(av[i + ky], wv[j], acc[i][j]);
                }
            }
        }
    };
    if constexpr (PF) {
        // Software pipeline (DenseLayer forward at the large levels): the raw activations and weights of chunk k+1 are
        // requested into registers before the FMA loop of chunk k and transformed / stored after it, so the memory
        // latency of a chunk is paid behind a full FMA phase.  BatchNorm coefficients of all K channels sit in shared memory.
        static_assert(!PF || (LM == LM_BNRELU && !UP && KCT == 4 && WM == WM_FWD), "prefetch variant: DenseLayer forward, 4-channel chunks");
        constexpr int NAP = (TRP * TWP + NT - 1) / NT, NWP = (CO * KCT * TAPS + NT - 1) / NT;
        float4* coef_s = reinterpret_cast<float4*>(w_s + KCT * TAPS * CO);           // [K] (a, beta, mean, invstd)
        for (int i = threadIdx.x; i < A.K; i += NT) coef_s[i] = ldg4(A.coef + ((size_t)g * A.K + i) * 4);
        float4 pa[NAP];
        float pw[NWP];
        unsigned okm = 0u;
        auto issue = [&](int k0) {
            okm = 0u;
#pragma unroll
            for (int j = 0; j < NAP; ++j) {
                const int pix = threadIdx.x + j * NT;
                const int r = pix / TWP, c = pix - r * TWP;
                const int y = y0 + r - PAD, x = x0 + c - PAD;
                pa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pix < TRP * TWP && y >= 0 && y < A.oh && x >= 0 && x < A.ow && k0 < A.K) {
                    pa[j] = ldg4(A.in + ((size_t)(b * A.ih + y) * A.iw + x) * A.in_C + A.in_off + k0);
                    okm |= 1u << j;
                }
            }
#pragma unroll
            for (int j = 0; j < NWP; ++j) {
                const int idx = threadIdx.x + j * NT;
                pw[j] = 0.f;
                if (idx < CO * KCT * TAPS) {
                    const int tap = idx % TAPS, kk = (idx / TAPS) % KCT, n = idx / (TAPS * KCT);
                    const int k = k0 + kk, nn = n0 + n;
                    if (k < A.K && nn < A.N) pw[j] = __ldg(A.w + ((size_t)nn * A.w_cin + k) * TAPS + tap);
                }
            }
        };
        auto commit = [&](int k0) {
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, c2 = c0, c3 = c0;
            if (k0 + 3 < A.K) { c0 = coef_s[k0]; c1 = coef_s[k0 + 1]; c2 = coef_s[k0 + 2]; c3 = coef_s[k0 + 3]; }
#pragma unroll
            for (int j = 0; j < NAP; ++j) {
                const int pix = threadIdx.x + j * NT;
                if (pix < TRP * TWP) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (okm & (1u << j)) {
                        v.x = fmaxf(fmaf(c0.x, pa[j].x - c0.z, c0.y), 0.f); v.y = fmaxf(fmaf(c1.x, pa[j].y - c1.z, c1.y), 0.f);
                        v.z = fmaxf(fmaf(c2.x, pa[j].z - c2.z, c2.y), 0.f); v.w = fmaxf(fmaf(c3.x, pa[j].w - c3.z, c3.y), 0.f);
                    }
                    float* d = a_s + pix;
                    d[0] = v.x; d[PLANE] = v.y; d[2 * PLANE] = v.z; d[3 * PLANE] = v.w;
                }
            }
#pragma unroll
            for (int j = 0; j < NWP; ++j) {
                const int idx = threadIdx.x + j * NT;
                if (idx < CO * KCT * TAPS) {
                    const int tap = idx % TAPS, kk = (idx / TAPS) % KCT, n = idx / (TAPS * KCT);
                    w_s[(kk * TAPS + tap) * CO + n] = pw[j];
                }
            }
        };
        if (k_lo < k_hi) issue(k_lo);
        for (int k0 = k_lo; k0 < k_hi; k0 += KCT) {
            __syncthreads();                 // FMA loop of the previous chunk is done with a_s / w_s (first pass: coef_s is written)
            commit(k0);
            __syncthreads();
            if (k0 + KCT < k_hi) issue(k0 + KCT);
            compute_chunk();
        }
    } else
    for (int k0 = k_lo; k0 < k_hi; k0 += KCT) {
        __syncthreads();
        // ---- stage operand A: global (NHWC, transformed on the fly) -> shared [k][row][col]
        if constexpr (LM == LM_NCHW) {
            for (int idx = threadIdx.x; idx < KCT * TRP * TWP; idx += NT) {
                const int kk = idx / (TRP * TWP), pix = idx - kk * (TRP * TWP);
                const int r = pix / TWP, c = pix - r * TWP;
                const int y = y0 + r - PAD, x = x0 + c - PAD, ch = k0 + kk;
                float v = 0.f;
                if (ch < A.K && y >= 0 && y < A.oh && x >= 0 && x < A.ow)
                    v = __ldg(A.in + ((size_t)(b * A.K + ch) * A.ih + y) * A.iw + x);
                a_s[kk * PLANE + pix] = v;
            }
        } else {
            // loads first, stores second (batches of 6): every thread keeps several 16-byte loads in flight instead
            // of one load -> transform -> store round trip per iteration
            constexpr int NA = (TRP * TWP * (KCT / 4) + NT - 1) / NT;
#pragma unroll
            for (int b0 = 0; b0 < NA; b0 += 6) {
                float4 v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b0 + j < NA && idx < TRP * TWP * (KCT / 4)) {
                        const int q = idx % (KCT / 4), pix = idx / (KCT / 4);
                        const int r = pix / TWP, c = pix - r * TWP;
                        v[j] = load_a4<LM, UP>(A, b, g, y0 + r - PAD, x0 + c - PAD, k0 + q * 4);
                    }
                }
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    if (b0 + j < NA && idx < TRP * TWP * (KCT / 4)) {
                        const int q = idx % (KCT / 4), pix = idx / (KCT / 4);
                        float* d = a_s + (q * 4) * PLANE + pix;
                        d[0] = v[j].x; d[PLANE] = v[j].y; d[2 * PLANE] = v[j].z; d[3 * PLANE] = v[j].w;
                    }
                }
            }
        }
        // ---- stage weights: w_s[kk][tap][n]  (all loads of this chunk issued before the first store)
        {
            constexpr int NWV = (CO * KCT * TAPS + NT - 1) / NT;
            float wv[NWV];
#pragma unroll
            for (int j = 0; j < NWV; ++j) {
                const int idx = threadIdx.x + j * NT;
                wv[j] = 0.f;
                if (idx < CO * KCT * TAPS) {
                    const int tap = idx % TAPS, kk = (idx / TAPS) % KCT, n = idx / (TAPS * KCT);
                    const int k = k0 + kk, nn = n0 + n;
                    if (k < A.K && nn < A.N) {
                        if constexpr (WM == WM_FWD) wv[j] = __ldg(A.w + ((size_t)nn * A.w_cin + k) * TAPS + tap);
                        else wv[j] = __ldg(A.w + ((size_t)k * A.w_cin + nn) * TAPS + (TAPS - 1 - tap));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NWV; ++j) {
                const int idx = threadIdx.x + j * NT;
                if (idx < CO * KCT * TAPS) {
                    const int tap = idx % TAPS, kk = (idx / TAPS) % KCT, n = idx / (TAPS * KCT);
                    w_s[(kk * TAPS + tap) * CO + n] = wv[j];
                }
            }
        }
        __syncthreads();
        compute_chunk();
    }

    // =================================================================================== epilogue
    __syncthreads();                             // smem is reused for the cross-warp reduction
    float* red = smem;                           // [NW][CO][2]
    const int x = x0 + lane;
    float s1[CO], s2[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) { s1[j] = 0.f; s2[j] = 0.f; }

    if constexpr (EM == EM_PARTIAL) {
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const int y = y0 + r0 + i;
            if (y < A.oh && x < A.ow) {
                float* pp = A.partial + ((((size_t)blockIdx.y * A.B + b) * A.oh + y) * A.ow + x) * CO;
#pragma unroll
                for (int j = 0; j < CO; j += 4)
                    *reinterpret_cast<float4*>(pp + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
            }
        }
        return;
    } else if constexpr (EM == EM_STORE) {
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const int y = y0 + r0 + i;
            const bool ok = (y < A.oh) && (x < A.ow);
            float* op = A.out + ((size_t)(b * A.oh + y) * A.ow + x) * A.out_C + A.out_off + n0;
#pragma unroll
            for (int j = 0; j < CO; j += 4) {
                if (n0 + j < A.N) {
                    const float4 bq = ldg4(A.bias + n0 + j);
                    float4 v = make_float4(acc[i][j] + bq.x, acc[i][j + 1] + bq.y, acc[i][j + 2] + bq.z, acc[i][j + 3] + bq.w);
                    if (ok) {
                        *reinterpret_cast<float4*>(op + j) = v;
                        s1[j] += v.x; s2[j] += v.x * v.x; s1[j + 1] += v.y; s2[j + 1] += v.y * v.y;
                        s1[j + 2] += v.z; s2[j + 2] += v.z * v.z; s1[j + 3] += v.w; s2[j + 3] += v.w * v.w;
                    }
                }
            }
        }
    } else if constexpr (EM == EM_POOL) {
        static_assert(EM != EM_POOL || (PX % 2 == 0), "pooling needs an even number of rows per thread");
        const int oh2 = A.oh >> 1, ow2 = A.ow >> 1;
#pragma unroll
        for (int i = 0; i < PX; i += 2) {
            const int y = y0 + r0 + i;
            const bool ok = (y < A.oh) && (x < A.ow) && ((lane & 1) == 0);
            const size_t pp = (size_t)(b * oh2 + (y >> 1)) * ow2 + (x >> 1);
#pragma unroll
            for (int j = 0; j < CO; j += 4) {
                float m4[4];
                unsigned am4 = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float bb = (n0 + j + e < A.N) ? __ldg(A.bias + n0 + j + e) : 0.f;
                    const float v00 = acc[i][j + e] + bb, v10 = acc[i + 1][j + e] + bb;
                    const float v01 = __shfl_xor_sync(0xffffffffu, v00, 1), v11 = __shfl_xor_sync(0xffffffffu, v10, 1);
                    float m = v00; unsigned am = 0;                              // ATen max_pool2d: (val > max) || isnan(val)
                    if (v01 > m || v01 != v01) { m = v01; am = 1; }
                    if (v10 > m || v10 != v10) { m = v10; am = 2; }
                    if (v11 > m || v11 != v11) { m = v11; am = 3; }
                    m4[e] = m; am4 |= am << (8 * e);
                }
                if (ok && n0 + j < A.N) {
                    *reinterpret_cast<float4*>(A.out + pp * A.out_C + A.out_off + n0 + j) = make_float4(m4[0], m4[1], m4[2], m4[3]);
                    *reinterpret_cast<unsigned*>(A.argmax_out + pp * A.N + n0 + j) = am4;
#pragma unroll
                    for (int e = 0; e < 4; ++e) { s1[j + e] += m4[e]; s2[j + e] += m4[e] * m4[e]; }
                }
            }
        }
    } else if constexpr (EM == EM_DGRAD_BN) {
        // channel quad outermost: the BN-backward sums of 4 channels are complete after PX pixels and are reduced across
        // the warp right away, so only 8 of them are live next to the accumulators (2 CTAs / SM need <= 128 registers)
#pragma unroll
        for (int j = 0; j < CO; j += 4) {
            float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
            if (n0 + j < A.N) {
                const float* cf = A.ep_coef + ((size_t)g * A.N + n0 + j) * 4;
                const float4 c0 = ldg4(cf), c1 = ldg4(cf + 4), c2 = ldg4(cf + 8), c3 = ldg4(cf + 12);
#pragma unroll
                for (int i = 0; i < PX; ++i) {
                    const int y = y0 + r0 + i;
                    if ((y < A.oh) && (x < A.ow)) {
                        const size_t o = ((size_t)(b * A.oh + y) * A.ow + x) * A.out_C + A.out_off + n0;
                        const float4 xq = ldg4(A.x + o + j);
                        float4 gq = *reinterpret_cast<const float4*>(A.out + o + j);
                        const float d0 = xq.x - c0.z, d1 = xq.y - c1.z, d2 = xq.z - c2.z, d3 = xq.w - c3.z;
                        const float g0 = fmaf(c0.x, d0, c0.y) > 0.f ? acc[i][j] : 0.f;
                        const float g1 = fmaf(c1.x, d1, c1.y) > 0.f ? acc[i][j + 1] : 0.f;
                        const float g2 = fmaf(c2.x, d2, c2.y) > 0.f ? acc[i][j + 2] : 0.f;
                        const float g3 = fmaf(c3.x, d3, c3.y) > 0.f ? acc[i][j + 3] : 0.f;
                        t1[0] += g0; t2[0] += g0 * (d0 * c0.w);
                        t1[1] += g1; t2[1] += g1 * (d1 * c1.w);
                        t1[2] += g2; t2[2] += g2 * (d2 * c2.w);
                        t1[3] += g3; t2[3] += g3 * (d3 * c3.w);
                        gq.x = fmaf(c0.x, g0, gq.x); gq.y = fmaf(c1.x, g1, gq.y);
                        gq.z = fmaf(c2.x, g2, gq.z); gq.w = fmaf(c3.x, g3, gq.w);
                        *reinterpret_cast<float4*>(A.out + o + j) = gq;
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a = t1[e], c = t2[e];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                if (lane == 0) { red[(warp * CO + j + e) * 2] = a; red[(warp * CO + j + e) * 2 + 1] = c; }
            }
        }
    } else if constexpr (EM == EM_DGRAD_UP) {
        static_assert(EM != EM_DGRAD_UP || (PX % 2 == 0), "2x2 sum needs an even number of rows per thread");
        const int oh2 = A.oh >> 1, ow2 = A.ow >> 1;
#pragma unroll
        for (int i = 0; i < PX; i += 2) {
            const int y = y0 + r0 + i;
            const bool ok = (y < A.oh) && (x < A.ow) && ((lane & 1) == 0);
            const size_t o = ((size_t)(b * oh2 + (y >> 1)) * ow2 + (x >> 1)) * A.out_C + A.out_off + n0;
#pragma unroll
            for (int j = 0; j < CO; j += 4) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = acc[i][j + e] + acc[i + 1][j + e];
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    v[e] = t;
                }
                if (ok && n0 + j < A.N) {
                    float4 gq = *reinterpret_cast<const float4*>(A.out + o + j);
                    gq.x += v[0]; gq.y += v[1]; gq.z += v[2]; gq.w += v[3];
                    *reinterpret_cast<float4*>(A.out + o + j) = gq;
                }
            }
        }
    }

    if constexpr (EM != EM_DGRAD_UP) {
        // per-channel sums: warp shuffle tree -> shared -> one fp64 atomic per channel per CTA
#pragma unroll
        for (int j = 0; j < (EM == EM_DGRAD_BN ? 0 : CO); ++j) {
            float a = s1[j], c = s2[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if (lane == 0) { red[(warp * CO + j) * 2] = a; red[(warp * CO + j) * 2 + 1] = c; }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < CO * 2; t += NT) {
            const int j = t >> 1, which = t & 1;
            if (n0 + j < A.N) {
                double s = 0.0;
#pragma unroll
                for (int wq = 0; wq < NW; ++wq) s += (double)red[(wq * CO + j) * 2 + which];
                const int ch = (EM == EM_DGRAD_BN) ? (n0 + j) : (A.out_off + n0 + j);
                atomicAdd(A.stats + ((size_t)g * A.stats_C + ch) * 2 + which, s);
            }
        }
    }
}

template <int KS, int PX, int CO, int NW, int KCT = KC>
constexpr size_t conv_smem_bytes() {
    using T = ConvTile<KS, PX, NW>;
    size_t main = sizeof(float) * (size_t)(KCT * T::PLANE + KCT * KS * KS * CO);
    size_t red = sizeof(float) * (size_t)(NW * CO * 2);
    return main > red ? main : red;
}

// Split-K finish: out[p][n] = bias[n] + sum_s partial[s][p][n] (fixed order), per-channel statistics of the result.
// thread = (pixel, channel quad); a block owns a contiguous pixel range of one statistic group.
template <int CO>
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ partial, const float* __restrict__ bias, float* __restrict__ out,
                     double* __restrict__ stats, int ksplit, long long pixels, int per_group, int N, int out_C, int out_off,
                     int stats_C) {
    pdl_enter();
    constexpr int NQ = CO / 4;
    __shared__ float red[8][CO][2];
    const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    const int q = threadIdx.x & 3;
    float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < NQ && q * 4 < N) bq = ldg4(bias + q * 4);
    for (int i = blockIdx.x * 64 + (threadIdx.x >> 2); i < per_group; i += gridDim.x * 64) {
        if (q < NQ && q * 4 < N) {
            const long long p = (long long)g * per_group + i;
            float4 v = bq;
            for (int sidx = 0; sidx < ksplit; ++sidx) {
                const float4 t = ldg4(partial + ((size_t)sidx * pixels + p) * CO + q * 4);
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            *reinterpret_cast<float4*>(out + (size_t)p * out_C + out_off + q * 4) = v;
            s1[0] += v.x; s2[0] += v.x * v.x; s1[1] += v.y; s2[1] += v.y * v.y;
            s1[2] += v.z; s2[2] += v.z * v.z; s1[3] += v.w; s2[3] += v.w * v.w;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
        }
    }
    if (lane < 4 && lane < NQ) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[warp][lane * 4 + e][0] = s1[e]; red[warp][lane * 4 + e][1] = s2[e]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CO * 2; t += 256) {
        const int j = t >> 1, which = t & 1;
        if (j < N) {
            double sum = 0.0;
#pragma unroll
            for (int wq = 0; wq < 8; ++wq) sum += (double)red[wq][j][which];
            atomicAdd(stats + ((size_t)g * stats_C + out_off + j) * 2 + which, sum);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][ky][kx] = sum_p g_out[p][co] * a[p + tap][ci]   (and db[co] = sum_p g_out[p][co])
// GEMM with M = ci*taps, N = co, K = pixels.  lane = ci inside a 32-channel chunk, warps = (ky, co group,
// pixel subgroup); a thread keeps KS x CW accumulators over the whole pixel range of its CTA and issues
// one atomicAdd per weight at the end.
// ---------------------------------------------------------------------------------------------------
struct WgradArgs {
    // activations (operand a)
    const float* a_in; const float* a_coef; int a_C, a_off, a_K, a_h, a_w;   // a_K = Cin of the layer
    // output gradient (operand g)
    const float* g_in; const float* g_x; const float* g_ab; const unsigned char* g_argmax;
    int g_C, g_off, g_K, g_h, g_w;                                          // g_K = Cout of the layer
    int oh, ow, B, G;
    float* dw; float* db; int w_cin;
    int tiles_per_cta, n_tiles;
};

template <int KS, int CW, int NCG, int NPS, int LMA, int LMG, bool UP>
__global__ void __launch_bounds__(KS * NCG * NPS * 32)
wgrad_kernel(const WgradArgs A) {
    pdl_enter();
    constexpr int PAD = KS / 2, TR = 8, TRP = TR + 2 * PAD, TWP = 32 + 2 * PAD, COT = CW * NCG;
    constexpr int NT = KS * NCG * NPS * 32;
    extern __shared__ __align__(16) float smem[];
    float* a_s = smem;                          // [TRP][TWP][32]   (ci fastest)
    float* g_s = smem + TRP * TWP * 32;         // [TR][32][COT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ky = warp % KS, cg = (warp / KS) % NCG, ps = warp / (KS * NCG);
    const int c0 = blockIdx.y * 32;             // ci chunk
    const int co0 = blockIdx.z * COT;           // co chunk
    const int tiles_x = (A.ow + 31) / 32, tiles_y = (A.oh + TR - 1) / TR;

    // ConvArgs views so the conv loaders can be reused
    ConvArgs LA{}; LA.in = A.a_in; LA.coef = A.a_coef; LA.in_C = A.a_C; LA.in_off = A.a_off; LA.K = A.a_K;
    LA.ih = A.a_h; LA.iw = A.a_w; LA.oh = A.oh; LA.ow = A.ow; LA.B = A.B; LA.G = A.G;
    ConvArgs LG{}; LG.in = A.g_in; LG.in2 = A.g_x; LG.in_ab = A.g_ab; LG.argmax = A.g_argmax; LG.in_C = A.g_C;
    LG.in_off = A.g_off; LG.K = A.g_K; LG.ih = A.g_h; LG.iw = A.g_w; LG.oh = A.oh; LG.ow = A.ow; LG.B = A.B; LG.G = A.G;

    float acc[KS][CW];
#pragma unroll
    for (int i = 0; i < KS; ++i)
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool do_bias = (blockIdx.y == 0) && (ky == 0) && (lane < CW);

    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    for (int t = t_begin; t < t_end; ++t) {
        const int b = t / (tiles_x * tiles_y), rem = t - b * (tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int y0 = ty * TR, x0 = tx * 32;
        const int g = b / (A.B / A.G);
        __syncthreads();
        // stage a: [row][col][ci]
        if constexpr (LMA == LM_NCHW) {
            for (int idx = threadIdx.x; idx < TRP * TWP * 32; idx += NT) {
                const int pix = idx % (TRP * TWP), ci = idx / (TRP * TWP);
                const int r = pix / TWP, c = pix - r * TWP;
                const int y = y0 + r - PAD, x = x0 + c - PAD, ch = c0 + ci;
                float v = 0.f;
                if (ch < A.a_K && y >= 0 && y < A.oh && x >= 0 && x < A.ow)
                    v = __ldg(A.a_in + ((size_t)(b * A.a_K + ch) * A.a_h + y) * A.a_w + x);
                a_s[pix * 32 + ci] = v;
            }
        } else {
            constexpr int NA = (TRP * TWP * 8 + NT - 1) / NT;       // loads first, stores second (batches of 7)
#pragma unroll
            for (int b0 = 0; b0 < NA; b0 += 7) {
                float4 v[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b0 + j < NA && idx < TRP * TWP * 8) {
                        const int q = idx & 7, pix = idx >> 3;
                        const int r = pix / TWP, c = pix - r * TWP;
                        v[j] = load_a4<LMA, UP>(LA, b, g, y0 + r - PAD, x0 + c - PAD, c0 + q * 4);
                    }
                }
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    if (b0 + j < NA && idx < TRP * TWP * 8) *reinterpret_cast<float4*>(a_s + (idx >> 3) * 32 + (idx & 7) * 4) = v[j];
                }
            }
        }
        // stage g: [row][col][co]
        {
            constexpr int NG = (TR * 32 * (COT / 4) + NT - 1) / NT;
#pragma unroll
            for (int b0 = 0; b0 < NG; b0 += 6) {
                float4 v[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b0 + j < NG && idx < TR * 32 * (COT / 4)) {
                        const int q = idx % (COT / 4), pix = idx / (COT / 4);
                        v[j] = load_a4<LMG, false>(LG, b, g, y0 + (pix >> 5), x0 + (pix & 31), co0 + q * 4);
                    }
                }
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int idx = threadIdx.x + (b0 + j) * NT;
                    if (b0 + j < NG && idx < TR * 32 * (COT / 4)) {
                        const int q = idx % (COT / 4), pix = idx / (COT / 4);
                        *reinterpret_cast<float4*>(g_s + pix * COT + q * 4) = v[j];
                    }
                }
            }
        }
        __syncthreads();
        for (int r = ps; r < TR; r += NPS) {
            const float* ar = a_s + ((r + ky) * TWP) * 32 + lane;
            const float* gr = g_s + (r * 32) * COT + cg * CW;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if constexpr (KS == 3) { a1 = ar[0]; a2 = ar[32]; }
#pragma unroll 4
            for (int xx = 0; xx < 32; ++xx) {
                if constexpr (KS == 3) { a0 = a1; a1 = a2; a2 = ar[(xx + 2) * 32]; }
                else a0 = ar[xx * 32];
                float gv[CW];
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(gr + xx * COT + j * 4);
                    gv[j * 4] = q.x; gv[j * 4 + 1] = q.y; gv[j * 4 + 2] = q.z; gv[j * 4 + 3] = q.w;
                }
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    acc[0][j] = fmaf(a0, gv[j], acc[0][j]);
                    if constexpr (KS == 3) { acc[1][j] = fmaf(a1, gv[j], acc[1][j]); acc[2][j] = fmaf(a2, gv[j], acc[2][j]); }
                }
                if (do_bias) bsum += gr[xx * COT + lane];
            }
        }
    }
    const int ci = c0 + lane;
    if (ci < A.a_K) {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const int co = co0 + cg * CW + j;
            if (co < A.g_K) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
                    atomicAdd(A.dw + ((size_t)co * A.w_cin + ci) * (KS * KS) + ky * KS + kx, acc[kx][j]);
            }
        }
    }
    if (do_bias && A.db && co0 + cg * CW + lane < A.g_K) atomicAdd(A.db + co0 + cg * CW + lane, bsum);
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient, software-pipelined version (same thread mapping and arithmetic as wgrad_kernel): the global
// loads of tile t+1 are issued into registers before the FMA loop of tile t and transformed / stored to shared
// memory after it, so a CTA pays the memory latency once instead of once per tile.  The BatchNorm coefficients
// and lazy-correction pairs of the CTA's channel chunk live in shared memory (loaded once).
// ---------------------------------------------------------------------------------------------------
template <int KS, int CW, int NCG, int NPS, int LMA, int LMG, bool UP>
__global__ void __launch_bounds__(KS * NCG * NPS * 32)
wgrad2_kernel(const WgradArgs A) {
    pdl_enter();
    static_assert(LMA == LM_BNRELU || LMA == LM_PLAIN, "activation loader");
    static_assert(LMG == LM_GRAD || LMG == LM_GRADPOOL, "gradient loader");
    constexpr int PAD = KS / 2, TR = 8, TRP = TR + 2 * PAD, TWP = 32 + 2 * PAD, COT = CW * NCG;
    constexpr int NT = KS * NCG * NPS * 32;
    constexpr int NA_ITEMS = TRP * TWP * 8, NG_ITEMS = TR * 32 * (COT / 4);
    constexpr int NA = (NA_ITEMS + NT - 1) / NT, NG = (NG_ITEMS + NT - 1) / NT;
    static_assert(NT % 8 == 0, "a thread must own one activation channel quad");
    extern __shared__ __align__(16) float smem[];
    float* a_s = smem;                                  // [TRP][TWP][32]   (ci fastest)
    float* g_s = smem + TRP * TWP * 32;                 // [TR][32][COT]
    float* coef_s = g_s + TR * 32 * COT;                // [2][32][4]  (a, beta, mean, invstd)
    float* ab_s = coef_s + 2 * 32 * 4;                  // [2][COT][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ky = warp % KS, cg = (warp / KS) % NCG, ps = warp / (KS * NCG);
    const int c0 = blockIdx.y * 32, co0 = blockIdx.z * COT;
    const int tiles_x = (A.ow + 31) / 32, tiles_y = (A.oh + TR - 1) / TR;
    const int per_img = tiles_x * tiles_y;

    for (int i = threadIdx.x; i < 2 * 32 * 4; i += NT) {
        const int gg = i / 128, ch = (i >> 2) & 31;
        float v = 0.f;
        if (LMA == LM_BNRELU && gg < A.G && c0 + ch < A.a_K) v = __ldg(A.a_coef + ((size_t)gg * A.a_K + c0 + ch) * 4 + (i & 3));
        coef_s[i] = v;
    }
    for (int i = threadIdx.x; i < 2 * COT * 2; i += NT) {
        const int gg = i / (COT * 2), ch = (i >> 1) % COT;
        float v = 0.f;
        if (gg < A.G && co0 + ch < A.g_K) v = __ldg(A.g_ab + ((size_t)gg * A.g_C + A.g_off + co0 + ch) * 2 + (i & 1));
        ab_s[i] = v;
    }

    const int qa = threadIdx.x & 7;                                         // this thread's activation channel quad
    float4 pa[NA], pg[NG], px[NG];
    unsigned pam[NG];
    unsigned ok_a = 0u, ok_g = 0u;

    auto issue = [&](int t) {
        const int b = t / per_img, rem = t - b * per_img;
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int y0 = ty * TR, x0 = tx * 32;
        ok_a = 0u; ok_g = 0u;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int idx = threadIdx.x + j * NT, pix = idx >> 3;
            const int r = pix / TWP, c = pix - r * TWP;
            const int y = y0 + r - PAD, x = x0 + c - PAD, ch = c0 + qa * 4;
            pa[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < NA_ITEMS && y >= 0 && y < A.oh && x >= 0 && x < A.ow && ch < A.a_K) {
                const int sy = UP ? (y >> 1) : y, sx = UP ? (x >> 1) : x;
                pa[j] = ldg4(A.a_in + ((size_t)(b * A.a_h + sy) * A.a_w + sx) * A.a_C + A.a_off + ch);
                ok_a |= 1u << j;
            }
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            const int idx = threadIdx.x + j * NT, pix = idx / (COT / 4), qg = idx - pix * (COT / 4);
            const int y = y0 + (pix >> 5), x = x0 + (pix & 31), ch = co0 + qg * 4;
            pg[j] = make_float4(0.f, 0.f, 0.f, 0.f); px[j] = pg[j]; pam[j] = 0xffffffffu;
            if (idx < NG_ITEMS && y < A.oh && x < A.ow && ch < A.g_K) {
                size_t pp;
                if constexpr (LMG == LM_GRADPOOL) {
                    pp = (size_t)(b * A.g_h + (y >> 1)) * A.g_w + (x >> 1);
                    pam[j] = __ldg(reinterpret_cast<const unsigned*>(A.g_argmax + pp * A.g_K + ch));
                } else {
                    pp = (size_t)(b * A.g_h + y) * A.g_w + x;
                }
                const size_t o = pp * A.g_C + A.g_off + ch;
                pg[j] = ldg4(A.g_in + o); px[j] = ldg4(A.g_x + o);
                ok_g |= 1u << j;
            }
        }
    };
    auto commit = [&](int t) {
        const int b = t / per_img, rem = t - b * per_img;
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int y0 = ty * TR, x0 = tx * 32;
        const int g = b / (A.B / A.G);
        float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0, k3 = k0;
        if constexpr (LMA == LM_BNRELU) {
            const float4* cf = reinterpret_cast<const float4*>(coef_s + (g * 32 + qa * 4) * 4);
            k0 = cf[0]; k1 = cf[1]; k2 = cf[2]; k3 = cf[3];
        }
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const int idx = threadIdx.x + j * NT;
            if (idx < NA_ITEMS) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok_a & (1u << j)) {
                    if constexpr (LMA == LM_BNRELU) {
                        v.x = fmaxf(fmaf(k0.x, pa[j].x - k0.z, k0.y), 0.f); v.y = fmaxf(fmaf(k1.x, pa[j].y - k1.z, k1.y), 0.f);
                        v.z = fmaxf(fmaf(k2.x, pa[j].z - k2.z, k2.y), 0.f); v.w = fmaxf(fmaf(k3.x, pa[j].w - k3.z, k3.y), 0.f);
                    } else {
                        v = pa[j];
                    }
                }
                *reinterpret_cast<float4*>(a_s + (idx >> 3) * 32 + qa * 4) = v;
            }
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            const int idx = threadIdx.x + j * NT;
            if (idx < NG_ITEMS) {
                const int pix = idx / (COT / 4), qg = idx - pix * (COT / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok_g & (1u << j)) {
                    const float4* abp = reinterpret_cast<const float4*>(ab_s + (g * COT + qg * 4) * 2);
                    const float4 c0q = abp[0], c1q = abp[1];
                    v.x = pg[j].x + fmaf(c0q.y, px[j].x, c0q.x); v.y = pg[j].y + fmaf(c0q.w, px[j].y, c0q.z);
                    v.z = pg[j].z + fmaf(c1q.y, px[j].z, c1q.x); v.w = pg[j].w + fmaf(c1q.w, px[j].w, c1q.z);
                    if constexpr (LMG == LM_GRADPOOL) {
                        const unsigned pos = (unsigned)((((y0 + (pix >> 5)) & 1) << 1) | ((x0 + (pix & 31)) & 1));
                        const unsigned am = pam[j];
                        if ((am & 0xffu) != pos) v.x = 0.f;
                        if (((am >> 8) & 0xffu) != pos) v.y = 0.f;
                        if (((am >> 16) & 0xffu) != pos) v.z = 0.f;
                        if ((am >> 24) != pos) v.w = 0.f;
                    }
                }
                *reinterpret_cast<float4*>(g_s + pix * COT + qg * 4) = v;
            }
        }
    };

    float acc[KS][CW];
#pragma unroll
    for (int i = 0; i < KS; ++i)
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;
    const bool do_bias = (blockIdx.y == 0) && (ky == 0) && (lane < CW);

    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    if (t_begin < t_end) issue(t_begin);
    for (int t = t_begin; t < t_end; ++t) {
        __syncthreads();                     // previous tile's FMA loop is done with the shared tiles (and coef_s / ab_s are written)
        commit(t);
        __syncthreads();
        if (t + 1 < t_end) issue(t + 1);
        for (int r = ps; r < TR; r += NPS) {
            const float* ar = a_s + ((r + ky) * TWP) * 32 + lane;
            const float* gr = g_s + (r * 32) * COT + cg * CW;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            if constexpr (KS == 3) { a1 = ar[0]; a2 = ar[32]; }
#pragma unroll 4
            for (int xx = 0; xx < 32; ++xx) {
                if constexpr (KS == 3) { a0 = a1; a1 = a2; a2 = ar[(xx + 2) * 32]; }
                else a0 = ar[xx * 32];
                float gv[CW];
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(gr + xx * COT + j * 4);
                    gv[j * 4] = q.x; gv[j * 4 + 1] = q.y; gv[j * 4 + 2] = q.z; gv[j * 4 + 3] = q.w;
                }
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    acc[0][j] = fmaf(a0, gv[j], acc[0][j]);
                    if constexpr (KS == 3) { acc[1][j] = fmaf(a1, gv[j], acc[1][j]); acc[2][j] = fmaf(a2, gv[j], acc[2][j]); }
                }
                if (do_bias) bsum += gr[xx * COT + lane];
            }
        }
    }
    const int ci = c0 + lane;
    if (ci < A.a_K) {
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const int co = co0 + cg * CW + j;
            if (co < A.g_K) {
#pragma unroll
                for (int kx = 0; kx < KS; ++kx)
                    atomicAdd(A.dw + ((size_t)co * A.w_cin + ci) * (KS * KS) + ky * KS + kx, acc[kx][j]);
            }
        }
    }
    if (do_bias && A.db && co0 + cg * CW + lane < A.g_K) atomicAdd(A.db + co0 + cg * CW + lane, bsum);
}

template <int KS, int CW, int NCG>
constexpr size_t wgrad2_smem_bytes() {
    constexpr int PAD = KS / 2;
    return sizeof(float) * (size_t)((8 + 2 * PAD) * (32 + 2 * PAD) * 32 + 8 * 32 * CW * NCG + 2 * 32 * 4 + 2 * CW * NCG * 2);
}

template <int KS, int CW, int NCG>
constexpr size_t wgrad_smem_bytes() {
    constexpr int PAD = KS / 2;
    return sizeof(float) * (size_t)((8 + 2 * PAD) * (32 + 2 * PAD) * 32 + 8 * 32 * CW * NCG);
}

// ---------------------------------------------------------------------------------------------------
// BatchNorm bookkeeping (tiny kernels, one thread per channel)
// ---------------------------------------------------------------------------------------------------
struct BnPrepArgs {
    const double* stats;      // [G][Ctot][2] of the level buffer
    float* mi;                // [G][Ctot][2] (mean, invstd)
    float* coef;              // [G][C][4] (a, beta, mean, invstd) for this BN
    const float* gamma; const float* beta;
    float* rmean; float* rvar;
    int C, Ctot, ch_off, G, training;
    double count;             // pixels per group
};

__global__ void bn_prepare_kernel(const BnPrepArgs A) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.C) return;
    const float gam = A.gamma[c], bet = A.beta[c];
    if (!A.training) {
        const float inv = 1.0f / sqrtf(A.rvar[c] + kBnEps);
        for (int g = 0; g < A.G; ++g) {
            float* cf = A.coef + ((size_t)g * A.C + c) * 4;
            cf[0] = gam * inv; cf[1] = bet; cf[2] = A.rmean[c]; cf[3] = inv;
            A.mi[((size_t)g * A.Ctot + A.ch_off + c) * 2] = A.rmean[c];
            A.mi[((size_t)g * A.Ctot + A.ch_off + c) * 2 + 1] = inv;
        }
        return;
    }
    float rm = A.rmean[c], rv = A.rvar[c];
    for (int g = 0; g < A.G; ++g) {
        const double s = A.stats[((size_t)g * A.Ctot + A.ch_off + c) * 2];
        const double q = A.stats[((size_t)g * A.Ctot + A.ch_off + c) * 2 + 1];
        const double mean = s / A.count;
        double var = q / A.count - mean * mean;
        if (var < 0.0) var = 0.0;
        const double inv = 1.0 / sqrt(var + (double)kBnEps);
        float* cf = A.coef + ((size_t)g * A.C + c) * 4;
        cf[0] = (float)((double)gam * inv); cf[1] = bet; cf[2] = (float)mean; cf[3] = (float)inv;
        A.mi[((size_t)g * A.Ctot + A.ch_off + c) * 2] = (float)mean;
        A.mi[((size_t)g * A.Ctot + A.ch_off + c) * 2 + 1] = (float)inv;
        // running buffers: momentum 0.1, unbiased variance (nn.BatchNorm2d); groups update in order
        const double unb = A.count > 1.0 ? var * (A.count / (A.count - 1.0)) : var;
        rm = (1.0f - kBnMomentum) * rm + kBnMomentum * (float)mean;
        rv = (1.0f - kBnMomentum) * rv + kBnMomentum * (float)unb;
    }
    A.rmean[c] = rm; A.rvar[c] = rv;
}

struct BnBwdArgs {
    double* red;              // [G][maxC][2] sums (sum gy, sum gy*xhat); zeroed again on exit
    int red_C;
    const float* coef;        // [G][C][4] (a, beta, mean, invstd)
    const float* mi;          // unused (kept for ABI stability of the struct)
    float* ab;                // [G][Ctot][2] lazy correction of the level buffer (accumulated)
    float* dgamma; float* dbeta;
    int C, Ctot, ch_off, G;
    double count;
};

__global__ void bn_bwd_finalize_kernel(const BnBwdArgs A) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.C) return;
    double dg = 0.0, db = 0.0;
    for (int g = 0; g < A.G; ++g) {
        double* r = A.red + ((size_t)g * A.red_C + c) * 2;
        const double s1 = r[0], s2 = r[1];
        r[0] = 0.0; r[1] = 0.0;
        dg += s2; db += s1;
        const float* cf = A.coef + ((size_t)g * A.C + c) * 4;
        const double a = cf[0], mean = cf[2], inv = cf[3];
        // dL/dx = a*gy - (a/N) * (S1 + xhat * S2),  xhat = (x - mean) * inv   ->  affine in x, applied lazily
        float* ab = A.ab + ((size_t)g * A.Ctot + A.ch_off + c) * 2;
        ab[0] += (float)(-(a / A.count) * (s1 - mean * inv * s2));
        ab[1] += (float)(-(a / A.count) * inv * s2);
    }
    A.dgamma[c] += (float)dg;
    A.dbeta[c] += (float)db;
}

// ---------------------------------------------------------------------------------------------------
// firstconv weight / bias gradient (models.py:111-113: conv3x3 3 -> 48 on the NCHW input images):
//   dW[co][ci][ky][kx] = sum_p G[p][co] * img[ci][p + (ky-1, kx-1)],   db[co] = sum_p G[p][co],   G = g + A_c + B_c x (lazy BN term)
// 1,296 outputs against 48 gradient channels per pixel: the generic kernel (32-channel input chunks, one of 3 used) spent
// 1.3 ms here.  A block stages a 4x32 gradient tile (24 KB) and the 3 x 6 x 34 image patch; thread = (pixel stream, 4
// output channels, 7 taps) keeps 28 accumulators in registers: one 16-byte + seven 4-byte shared loads per 28 FMAs.
// ---------------------------------------------------------------------------------------------------
constexpr int FW_STREAMS = 5, FW_THREADS = FW_STREAMS * 48, FW_ROWS = 4, FW_PX = FW_ROWS * 32, FW_PR = FW_ROWS + 2;
__global__ void __launch_bounds__(FW_THREADS)
first_wgrad_kernel(const WgradArgs A) {
    pdl_enter();
    __shared__ __align__(16) float g_s[FW_PX * 48];
    __shared__ float in_s[3 * FW_PR * 34];
    const int tid = threadIdx.x;
    const int stream = tid / 48, o = tid - stream * 48;
    const int cq = o % 12, tg = o / 12;                  // output-channel quad, tap group (taps 7 tg .. 7 tg + 6 of 27)
    int tap_off[7];                                      // offset of tap k inside the patch, relative to the pixel's (row, col)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int k = tg * 7 + j, kk = k < 27 ? k : 26;
        const int ci = kk / 9, ky = (kk % 9) / 3, kx = kk % 3;
        tap_off[j] = (ci * FW_PR + ky) * 34 + kx;
    }
    float acc[4][7], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[e][j] = 0.f;
    const int tiles_x = (A.ow + 31) / 32, tiles_y = (A.oh + FW_ROWS - 1) / FW_ROWS;
    const int n_tiles = A.B * tiles_x * tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int b = t / (tiles_x * tiles_y), rem = t - b * (tiles_x * tiles_y);
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        const int y0 = ty * FW_ROWS, x0 = tx * 32;
        const int g = b / (A.B / A.G);
        __syncthreads();                                 // previous tile fully consumed
        // gradient tile [256 px][48] with the lazy BN correction
        for (int i = tid; i < FW_PX * 12; i += FW_THREADS) {
            const int px = i / 12, q = i - px * 12;
            const int y = y0 + (px >> 5), x = x0 + (px & 31);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < A.oh && x < A.ow) {
                const size_t off = ((size_t)(b * A.oh + y) * A.ow + x) * A.g_C + A.g_off + q * 4;
                const float4 gq = ldg4(A.g_in + off), xq = ldg4(A.g_x + off);
                const float* abp = A.g_ab + ((size_t)g * A.g_C + A.g_off + q * 4) * 2;
                const float4 c0 = ldg4(abp), c1 = ldg4(abp + 4);
                v.x = gq.x + fmaf(c0.y, xq.x, c0.x); v.y = gq.y + fmaf(c0.w, xq.y, c0.z);
                v.z = gq.z + fmaf(c1.y, xq.z, c1.x); v.w = gq.w + fmaf(c1.w, xq.w, c1.z);
            }
            *reinterpret_cast<float4*>(g_s + px * 48 + q * 4) = v;
        }
        // image patch (NCHW, zero padding)
        for (int i = tid; i < 3 * FW_PR * 34; i += FW_THREADS) {
            const int ci = i / (FW_PR * 34), r = (i - ci * (FW_PR * 34)) / 34, cc = i - ci * (FW_PR * 34) - r * 34;
            const int y = y0 + r - 1, x = x0 + cc - 1;
            float v = 0.f;
            if (y >= 0 && y < A.a_h && x >= 0 && x < A.a_w) v = __ldg(A.a_in + ((size_t)(b * 3 + ci) * A.a_h + y) * A.a_w + x);
            in_s[i] = v;
        }
        __syncthreads();
        for (int px = stream; px < FW_PX; px += FW_STREAMS) {
            const float4 g4 = *reinterpret_cast<const float4*>(g_s + px * 48 + cq * 4);
            const float* ip = in_s + (px >> 5) * 34 + (px & 31);
            if (tg == 0) { bsum[0] += g4.x; bsum[1] += g4.y; bsum[2] += g4.z; bsum[3] += g4.w; }
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float v = ip[tap_off[j]];
                acc[0][j] = fmaf(g4.x, v, acc[0][j]); acc[1][j] = fmaf(g4.y, v, acc[1][j]);
                acc[2][j] = fmaf(g4.z, v, acc[2][j]); acc[3][j] = fmaf(g4.w, v, acc[3][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int k = tg * 7 + j;
        if (k < 27) {
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(A.dw + (size_t)(cq * 4 + e) * 27 + k, acc[e][j]);
        }
    }
    if (tg == 0 && A.db) {
#pragma unroll
        for (int e = 0; e < 4; ++e) atomicAdd(A.db + cq * 4 + e, bsum[e]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Conv bias gradient alone: db[co] = sum_p (g + A + B x)[p][co].  Only used when the weight gradient runs on the
// tensor cores but the data gradient does not (debug A/B combinations); normally the dgrad kernel produces it.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ ab, float* __restrict__ db,
                 int C, int off, int Cout, long long npix_per_group, int G) {
    pdl_enter();
    // blockIdx.y selects a group of 16 output channels
    off += 16 * blockIdx.y; db += 16 * blockIdx.y; Cout -= 16 * blockIdx.y;
    if (Cout > 16) Cout = 16;
    __shared__ float red[8][16];
    const int grp = threadIdx.x & 3, ch = grp * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < Cout) {
        for (int gi = 0; gi < G; ++gi) {
            const float* abp = ab + ((size_t)gi * C + off + ch) * 2;
            const float4 c0 = ldg4(abp), c1 = ldg4(abp + 4);
            for (long long p = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); p < npix_per_group; p += (long long)gridDim.x * 64) {
                const size_t o = ((size_t)gi * npix_per_group + p) * C + off + ch;
                const float4 gq = ldg4(g + o), xq = ldg4(x + o);
                s.x += gq.x + fmaf(c0.y, xq.x, c0.x); s.y += gq.y + fmaf(c0.w, xq.y, c0.z);
                s.z += gq.z + fmaf(c1.y, xq.z, c1.x); s.w += gq.w + fmaf(c1.w, xq.w, c1.z);
            }
        }
    }
#pragma unroll
    for (int o2 = 4; o2 < 32; o2 <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o2); s.y += __shfl_xor_sync(0xffffffffu, s.y, o2);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o2); s.w += __shfl_xor_sync(0xffffffffu, s.w, o2);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 4) { red[warp][lane * 4] = s.x; red[warp][lane * 4 + 1] = s.y; red[warp][lane * 4 + 2] = s.z; red[warp][lane * 4 + 3] = s.w; }
    __syncthreads();
    if (threadIdx.x < 16 && threadIdx.x < Cout) {
        float t = 0.f;
        for (int wq = 0; wq < 8; ++wq) t += red[wq][threadIdx.x];
        atomicAdd(db + threadIdx.x, t);
    }
}

// ---------------------------------------------------------------------------------------------------
// finalConv 1x1 C -> 1 followed by abs (models.py:167-169, 186): 8 lanes per pixel, HBM-bound
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
final_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ pre, float* __restrict__ y, long long npix, int C) {
    pdl_enter();
    const int sub = threadIdx.x & 7;
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    float s = 0.f;
    if (p < npix) {
        const float* xp = x + p * C;
        for (int c = sub * 4; c < C; c += 32) {
            const float4 a = ldg4(xp + c), q = ldg4(w + c);
            s = fmaf(a.x, q.x, s); s = fmaf(a.y, q.y, s); s = fmaf(a.z, q.z, s); s = fmaf(a.w, q.w, s);
        }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (p < npix && sub == 0) {
        s += bias[0];
        pre[p] = s;
        y[p] = fabsf(s);
    }
}

// backward: g_pre = g_y * sign(pre); gx[p][c] = g_pre * w[c] (first writer of the level-0 gradient buffer);
// dW[c] += sum_p g_pre * x[p][c]; db += sum_p g_pre
__global__ void __launch_bounds__(256)
final_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ pre, const float* __restrict__ x,
                 const float* __restrict__ w, float* __restrict__ gx, float* __restrict__ dw, float* __restrict__ db,
                 long long npix, int C, int pix_per_cta) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];     // [C] dW partial
    for (int c = threadIdx.x; c < C; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;          // 32 pixel groups per CTA
    const long long p_begin = (long long)blockIdx.x * pix_per_cta;
    const long long p_end = min(p_begin + pix_per_cta, npix);
    float bsum = 0.f;
    // each 8-lane group walks pixels; lane `sub` owns channels sub*4 + 32*k
    constexpr int MAXQ = 12;                                            // supports C <= 384
    float wacc[MAXQ][4];
#pragma unroll
    for (int k = 0; k < MAXQ; ++k) { wacc[k][0] = wacc[k][1] = wacc[k][2] = wacc[k][3] = 0.f; }
    for (long long p = p_begin + grp; p < p_end; p += 32) {
        const float s = pre[p];
        const float gp = gy[p] * (float)((s > 0.f) - (s < 0.f));
        if (sub == 0) bsum += gp;
        const float* xp = x + p * C;
        float* gp_out = gx + p * C;
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int c = sub * 4 + 32 * k;
            if (c < C) {
                const float4 a = ldg4(xp + c), q = ldg4(w + c);
                *reinterpret_cast<float4*>(gp_out + c) = make_float4(gp * q.x, gp * q.y, gp * q.z, gp * q.w);
                wacc[k][0] = fmaf(gp, a.x, wacc[k][0]); wacc[k][1] = fmaf(gp, a.y, wacc[k][1]);
                wacc[k][2] = fmaf(gp, a.z, wacc[k][2]); wacc[k][3] = fmaf(gp, a.w, wacc[k][3]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MAXQ; ++k) {
        const int c = sub * 4 + 32 * k;
        if (c < C) {
#pragma unroll
            for (int e = 0; e < 4; ++e) atomicAdd(sm + c + e, wacc[k][e]);
        }
    }
    // bias: reduce over the CTA
    __shared__ float sb;
    if (threadIdx.x == 0) sb = 0.f;
    __syncthreads();
    if (sub == 0 && bsum != 0.f) atomicAdd(&sb, bsum);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dw + c, sm[c]);
    if (threadIdx.x == 0) atomicAdd(db, sb);
}

}  // namespace endo
