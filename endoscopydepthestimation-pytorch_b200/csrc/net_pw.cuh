// Pointwise (1x1) convolution of the TransitionDown layers on tcgen05 (reference models.py:56-67: BN -> ReLU ->
// conv1x1 C -> C -> MaxPool2d(2)), forward and data gradient, as ONE GEMM per 128-pixel tile over ALL output channels:
//
//   forward : T[p][co] = bias[co] + sum_ci relu(bn(x[p][ci])) * W[co][ci]            (T = scratch tensor; the 2x2 max-pool,
//                                                                                     argmax and statistics are td_pool_kernel's)
//   dgrad   : gx[p][ci] += a_ci * [bn(x)[p][ci] > 0] * sum_co R[p][co] * W[co][ci]    R = pooled gradient routed to the
//                                                                                     window position the forward pool chose
//
// M = 128 consecutive pixels of the NHWC buffer (TMEM lanes), N = C (96 .. 288: the whole accumulator row block, <= 512 TMEM
// columns, stays resident), K = C in chunks of 16 channels (8 in 3xTF32 mode).  The input is read exactly once (the
// previous version ran the 3x3 kernel in 1x1 mode, 48 output channels per pass: C/48 passes over the input; the data
// gradient ran on the FFMA pipe).  Operands are K-major SWIZZLE_NONE planes (16-byte chunk = 4 tf32 channels): plane
// stride = LBO, 8-row groups 128 B apart (SBO).  Four shared-memory stages, mbarrier full/empty ring; the activation
// loads of the next four chunks are in flight in registers while a chunk is transformed (BN+ReLU / routing, tf32
// rounding or hi/lo split) and stored; weights come as ready-made stage images (pack_w_pw_kernel) from L2.
//
// Warp roles (288 threads): warps 0-7 stage operands and run the epilogue, warp 8 lane 0 issues the MMAs.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace endo {
namespace tcpw {

using tcconv::tf32_rn;

constexpr int MT = 128;                        // pixels per CTA
constexpr int NST = 4;                         // pipeline stages
constexpr int A_STAGE = 4 * MT * 16;           // 4 planes x 128 rows x 16 B
constexpr int NPF = 4;                         // chunks of activation loads kept in flight per thread
constexpr int NTHREADS = 288;
constexpr int TB_PITCH = 64 * 4 + 16;          // transposed epilogue tile: 128 pixels x 64 channels (+ bank spread)
constexpr int TB_BYTES = MT * TB_PITCH;

// Stage image of the weights: chunk c -> [plane][n][4 floats]; n = GEMM N index (output channel in the forward, input
// channel in the data gradient), plane p = 4 consecutive GEMM-K channels.  x3 = 0: 16 K-channels per chunk, rounded to
// tf32; x3 = 1: 8 K-channels per chunk, planes 0,1 = hi (tf32-rounded), planes 2,3 = lo (exact remainder).
// transposed = 0: B[n][k] = W[n][k] (forward, OIHW with 1x1 taps); 1: B[n][k] = W[k][n] (data gradient).
__global__ void __launch_bounds__(256)
pack_w_pw_kernel(const float* __restrict__ w, int C, int Npad, int x3, int transposed, float* __restrict__ out) {
    pdl_enter();
    const int c = blockIdx.x;
    const int total = 4 * Npad * 4;
    for (int d = threadIdx.x; d < total; d += 256) {
        const int plane = d / (Npad * 4), r = d - plane * Npad * 4, n = r >> 2, e = r & 3;
        const int k = x3 ? (c * 8 + (plane & 1) * 4 + e) : (c * 16 + plane * 4 + e);
        float v = 0.f;
        if (n < C && k < C) v = transposed ? __ldg(w + (size_t)k * C + n) : __ldg(w + (size_t)n * C + k);
        const float hi = tf32_rn(v);
        out[(size_t)c * total + d] = (x3 && plane >= 2) ? (v - hi) : hi;
    }
}

struct Args {
    // ---- operand A, forward: activation buffer
    const float* in; int in_C, in_off;
    const float* coef;                         // [G][K][4] (a, beta, mean, invstd) of the TransitionDown BatchNorm
    // ---- operand A, data gradient: pooled gradient of the next level, routed by the forward argmax
    const unsigned char* argmax;               // [B, H/2, W/2, K]
    const float* gc; const float* xc; const float* abc;   // coarse gradient / activation buffers (stride cC, first channel c_off),
    int cC, c_off;                             // lazy BN correction [G][cC][2]
    int H, W;                                  // fine resolution
    const float* wpack;
    const float* bias;                         // forward
    float* out; int out_C, out_off;            // forward: scratch [P][out_C]; dgrad: gradient buffer, accumulated at out_off
    const float* x;                            // dgrad epilogue: activation buffer of this level (stride out_C, channels at out_off)
    const float* ep_coef;                      // dgrad epilogue: [G][N][4]
    double* red; int red_C;                    // dgrad epilogue: [G][red_C][2] BN-backward sums
    int K, N, Npad;
    long long per_group;                       // pixels per statistic group (grid = tiles per group x groups)
    int mode;                                  // 0 forward, 1 data gradient, 2 forward with the 2x2 max-pool fused into the epilogue:
                                               // M rows = 32 pool windows x 4 positions, per_group counts COARSE pixels, out = next
                                               // level buffer, argmax_out = window positions, red = its statistics [G][red_C][2]
    unsigned char* argmax_out;
    // mode 1 by-products (nullptr = off): bf16 [pixels][K] routed gradient (operand A as staged) and bf16 [pixels][N] relu(bn(x))
    // (what the epilogue evaluates for the ReLU mask): the operands of the weight-gradient GEMM (net_pwwgrad.cuh)
    unsigned short* r16; unsigned short* a16;
    int x3;
};

__host__ __device__ inline int b_stage_bytes(int Npad) { return 4 * Npad * 16; }
__host__ inline size_t smem_bytes(int Npad, int K, int mode) {
    size_t s = (size_t)NST * (A_STAGE + b_stage_bytes(Npad));
    s += (size_t)K * 16;                       // coefficient table of operand A ([K][4] forward, [K][2] dgrad)
    s += (size_t)Npad * 16;                    // bias (forward) / epilogue BN table [N][4] (dgrad)
    if (mode == 1) s += TB_BYTES + 8 * 64 * 2 * 4;
    s += 256;                                  // barriers + TMEM slot
    return s;
}

__global__ void __launch_bounds__(NTHREADS, 2)
pw_gemm_kernel(const Args A) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int bstage = b_stage_bytes(A.Npad);
    unsigned char* a_st = smem;
    unsigned char* b_st = smem + NST * A_STAGE;
    float* ktab = reinterpret_cast<float*>(b_st + NST * bstage);          // operand-A coefficients
    float* ntab = ktab + A.K * 4;                                          // bias / epilogue table
    unsigned char* tb = reinterpret_cast<unsigned char*>(ntab + A.Npad * 4);
    float* red = reinterpret_cast<float*>(tb + (A.mode == 1 ? TB_BYTES : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + (A.mode == 1 ? 8 * 64 * 2 : 0));   // full[NST], empty[NST], accum
    bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 15) & ~(uintptr_t)15);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y;                                 // statistic group: tiles never straddle two groups
    const long long p0 = (long long)g * A.per_group + (long long)blockIdx.x * (A.mode == 2 ? MT / 4 : MT);   // mode 2: coarse pixels
    const long long p_end = (long long)(g + 1) * A.per_group;
    const int kch = A.x3 ? 8 : 16;
    const int nchunks = (A.K + kch - 1) / kch;
    const uint32_t ncols = A.Npad <= 128 ? 128u : (A.Npad <= 256 ? 256u : 512u);

    pdl_trigger();
    if (warp == 8) tc::tmem_alloc(tmem_slot, ncols);
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(bars + i, 256); tc::mbar_init(bars + NST + i, 1); }
        tc::mbar_init(bars + 2 * NST, 1);
        tc::fence_mbar_init();
    }
    pdl_wait();
    if (warp < 8) {
        // tables: operand-A coefficients and the bias / epilogue BatchNorm table
        if (A.mode != 1) {
            for (int i = tid; i < A.K; i += 256)
                *reinterpret_cast<float4*>(ktab + i * 4) = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)g * A.K + i) * 4));
            for (int i = tid; i < A.Npad; i += 256) ntab[i] = (i < A.N) ? __ldg(A.bias + i) : 0.f;
        } else {
            for (int i = tid; i < A.K; i += 256) {
                const float2 ab = __ldg(reinterpret_cast<const float2*>(A.abc + ((size_t)g * A.cC + A.c_off + i) * 2));
                ktab[i * 2] = ab.x; ktab[i * 2 + 1] = ab.y;
            }
            for (int i = tid; i < A.Npad; i += 256) {
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < A.N) e = __ldg(reinterpret_cast<const float4*>(A.ep_coef + ((size_t)g * A.N + i) * 4));
                *reinterpret_cast<float4*>(ntab + i * 4) = e;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ======================================================================== producers
        // item = (pixel, 4-channel group of the chunk): tf32: 2 items per thread (4 groups), 3xTF32: 1 item (2 groups, hi + lo)
        const int nitem = A.x3 ? 1 : 2;
        int ipx[2], igrp[2];
        bool iok[2];
        size_t ioff[2];                                    // element offset of the item's pixel in the source buffer
        unsigned ipos[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = tid + 256 * j;
            ipx[j] = A.x3 ? (tid >> 1) : (i >> 2);
            igrp[j] = A.x3 ? (tid & 1) : (i & 3);
            const long long p = p0 + (A.mode == 2 ? (ipx[j] >> 2) : ipx[j]);
            iok[j] = p < p_end && j < nitem;
            ioff[j] = 0; ipos[j] = 0;
            if (iok[j]) {
                if (A.mode == 0) ioff[j] = (size_t)p * A.in_C + A.in_off + igrp[j] * 4;
                else if (A.mode == 2) {
                    // row = 4 * window + position: the four pixels of a pool window sit in four consecutive TMEM lanes
                    const int wc = A.W >> 1, hc = A.H >> 1;
                    const int x2 = (int)(p % wc), y2 = (int)((p / wc) % hc);
                    const long long bq = p / ((long long)wc * hc);
                    const int yq = 2 * y2 + ((ipx[j] >> 1) & 1), xq = 2 * x2 + (ipx[j] & 1);
                    ioff[j] = (size_t)((bq * A.H + yq) * A.W + xq) * A.in_C + A.in_off + igrp[j] * 4;
                } else {
                    const int xq = (int)(p % A.W), yq = (int)((p / A.W) % A.H);
                    const long long bq = p / ((long long)A.W * A.H);
                    ioff[j] = (size_t)((bq * (A.H >> 1) + (yq >> 1)) * (A.W >> 1) + (xq >> 1));   // coarse pixel index
                    ipos[j] = (unsigned)(((yq & 1) << 1) | (xq & 1));
                }
            }
        }
        // register ring: [NPF][item]; dgrad needs (argmax word, g, x) per item -> ring depth 2 there
        float4 qa[NPF][2];
        float4 qx[2][2];
        unsigned qm[2][2];
        auto issue = [&](int c, int slot) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j < nitem) {
                    const int ch = c * kch + igrp[j] * 4;
                    const bool ok = iok[j] && c < nchunks && ch < A.K;
                    if (A.mode != 1) {
                        qa[slot][j] = ok ? __ldg(reinterpret_cast<const float4*>(A.in + ioff[j] + c * kch)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        qa[slot & 1][j] = make_float4(0.f, 0.f, 0.f, 0.f); qx[slot & 1][j] = qa[slot & 1][j]; qm[slot & 1][j] = 0xffffffffu;
                        if (ok) {
                            qm[slot & 1][j] = __ldg(reinterpret_cast<const unsigned*>(A.argmax + ioff[j] * A.K + ch));
                            qa[slot & 1][j] = __ldg(reinterpret_cast<const float4*>(A.gc + ioff[j] * A.cC + A.c_off + ch));
                            qx[slot & 1][j] = __ldg(reinterpret_cast<const float4*>(A.xc + ioff[j] * A.cC + A.c_off + ch));
                        }
                    }
                }
            }
        };
        const int depth = A.mode != 1 ? NPF : 2;
#pragma unroll
        for (int u = 0; u < NPF; ++u)
            if (u < depth) issue(u, u);
        const int nb4 = A.Npad * 4;                         // float4 per weight stage image
        for (int c0 = 0; c0 < nchunks; c0 += NPF) {
#pragma unroll
            for (int u = 0; u < NPF; ++u) {
                const int c = c0 + u;
                if (c < nchunks) {
                    const int s = c % NST;
                    if (c >= NST) tc::mbar_wait(bars + NST + s, ((c / NST) - 1) & 1);
                    // weights of this chunk (L2-resident image): cp.async straight into the stage (no registers), in flight
                    // while the activations are transformed and stored
                    {
                        const float4* wsrc = reinterpret_cast<const float4*>(A.wpack) + (size_t)c * nb4;
                        unsigned char* b_s = b_st + s * bstage;
                        for (int i = tid; i < nb4; i += 256) tc::cp_async16(b_s + (size_t)i * 16, wsrc + i, 16u);
                        tc::cp_async_commit();
                    }
                    unsigned char* a_s = a_st + s * A_STAGE;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j < nitem) {
                            const int ch = c * kch + igrp[j] * 4;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (iok[j] && ch < A.K) {
                                if (A.mode != 1) {
                                    const float4 q = qa[u][j];
                                    const float4 k0 = *reinterpret_cast<const float4*>(ktab + (ch + 0) * 4), k1 = *reinterpret_cast<const float4*>(ktab + (ch + 1) * 4);
                                    const float4 k2 = *reinterpret_cast<const float4*>(ktab + (ch + 2) * 4), k3 = *reinterpret_cast<const float4*>(ktab + (ch + 3) * 4);
                                    v.x = fmaxf(fmaf(k0.x, q.x - k0.z, k0.y), 0.f); v.y = fmaxf(fmaf(k1.x, q.y - k1.z, k1.y), 0.f);
                                    v.z = fmaxf(fmaf(k2.x, q.z - k2.z, k2.y), 0.f); v.w = fmaxf(fmaf(k3.x, q.w - k3.z, k3.y), 0.f);
                                } else {
                                    const float4 gq = qa[u & 1][j], xq = qx[u & 1][j];
                                    const unsigned am = qm[u & 1][j], pos = ipos[j];
                                    const float4 c0f = *reinterpret_cast<const float4*>(ktab + ch * 2), c1f = *reinterpret_cast<const float4*>(ktab + ch * 2 + 4);
                                    v.x = ((am & 0xffu) == pos) ? gq.x + fmaf(c0f.y, xq.x, c0f.x) : 0.f;
                                    v.y = (((am >> 8) & 0xffu) == pos) ? gq.y + fmaf(c0f.w, xq.y, c0f.z) : 0.f;
                                    v.z = (((am >> 16) & 0xffu) == pos) ? gq.z + fmaf(c1f.y, xq.z, c1f.x) : 0.f;
                                    v.w = ((am >> 24) == pos) ? gq.w + fmaf(c1f.w, xq.w, c1f.z) : 0.f;
                                    if (A.r16)
                                        *reinterpret_cast<uint2*>(A.r16 + (size_t)(p0 + ipx[j]) * A.K + ch) =
                                            make_uint2(tcwgrad::pack_bf16(v.x, v.y), tcwgrad::pack_bf16(v.z, v.w));
                                }
                            }
                            const float4 hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                            *reinterpret_cast<float4*>(a_s + igrp[j] * (MT * 16) + ipx[j] * 16) = hi;
                            if (A.x3)
                                *reinterpret_cast<float4*>(a_s + (2 + igrp[j]) * (MT * 16) + ipx[j] * 16) =
                                    make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                        }
                    }
                    // refill the ring slot just consumed
                    if (A.mode != 1) issue(c + NPF, u); else issue(c + 2, u);
                    tc::cp_async_wait<0>();
                    tc::fence_proxy_async();
                    tc::mbar_arrive(bars + s);
                }
            }
        }
        // ======================================================================== epilogue
        tc::mbar_wait(bars + 2 * NST, 0);
        tc::tc_fence_after();
        const int q = warp & 3, hf = warp >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        if (A.mode == 0) {
            // 64-channel column blocks: TMEM -> transposed shared tile (the pipeline stages are idle now) -> lane = (pixel
            // parity, channel quad), so that a warp store covers two pixels x 256 contiguous bytes.  (One thread per pixel
            // writing 16-byte pieces 768 B apart cost 32 line transactions per store instruction: the store path, not HBM,
            // bounded this kernel.)
            unsigned char* tbf = smem;
            const int quad = lane & 15, psub = lane >> 4;
            for (int cb = 0; cb * 64 < A.Npad; ++cb) {
                {
                    const uint32_t taddr = tmem + lane_base + cb * 64 + hf * 32;
                    unsigned char* row = tbf + (size_t)(q * 32 + lane) * TB_PITCH + hf * 128;
                    float v[16];
                    if (cb * 64 + hf * 32 < A.Npad) {
                        tc::tmem_ld16(taddr, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (cb * 64 + hf * 32 + 16 < A.Npad) {
                        tc::tmem_ld16(taddr + 16, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + 64 + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int n0 = cb * 64 + quad * 4;
                if (n0 < A.N) {
                    const float4 bq = *reinterpret_cast<const float4*>(ntab + n0);
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int pl = warp * 16 + it * 2 + psub;
                        const long long p = p0 + pl;
                        if (p < p_end) {
                            const float4 d = *reinterpret_cast<const float4*>(tbf + (size_t)pl * TB_PITCH + quad * 16);
                            *reinterpret_cast<float4*>(A.out + (size_t)p * A.out_C + A.out_off + n0) =
                                make_float4(d.x + bq.x, d.y + bq.y, d.z + bq.z, d.w + bq.w);
                        }
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");           // tile reusable
            }
        } else if (A.mode == 2) {
            // forward with the max-pool fused: 64-channel column blocks, TMEM -> transposed shared tile (rows = 4 window + position)
            // -> thread = (window, channel quad): bias, max / argmax over the four rows in ATen's scan order, pooled value and
            // argmax word to the next level, statistics of the pooled map.  The full-resolution conv output never reaches HBM
            // (the separate td_pool pass wrote and re-read it: 2 x 503 MB at the first TransitionDown of a 256x320 batch of 16).
            unsigned char* tbf = smem;
            float* red2 = reinterpret_cast<float*>(smem + TB_BYTES);          // [8 warps][64][2], behind the tile (the stages are idle)
            const int quad = lane & 15, psub = lane >> 4;
            for (int cb = 0; cb * 64 < A.Npad; ++cb) {
                {
                    const uint32_t taddr = tmem + lane_base + cb * 64 + hf * 32;
                    unsigned char* row = tbf + (size_t)(q * 32 + lane) * TB_PITCH + hf * 128;
                    float v[16];
                    if (cb * 64 + hf * 32 < A.Npad) {
                        tc::tmem_ld16(taddr, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (cb * 64 + hf * 32 + 16 < A.Npad) {
                        tc::tmem_ld16(taddr + 16, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + 64 + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int n0 = cb * 64 + quad * 4;
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                if (n0 < A.N) {
                    const float4 bq = *reinterpret_cast<const float4*>(ntab + n0);
                    const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int wl = warp * 4 + it * 2 + psub;               // pool window of this tile
                        const long long cq = p0 + wl;                          // coarse pixel
                        if (cq < p_end) {
                            const unsigned char* r0 = tbf + (size_t)(4 * wl) * TB_PITCH + quad * 16;
                            const float4 d0 = *reinterpret_cast<const float4*>(r0), d1 = *reinterpret_cast<const float4*>(r0 + TB_PITCH);
                            const float4 d2 = *reinterpret_cast<const float4*>(r0 + 2 * TB_PITCH), d3 = *reinterpret_cast<const float4*>(r0 + 3 * TB_PITCH);
                            const float a0[4] = {d0.x, d0.y, d0.z, d0.w}, a1[4] = {d1.x, d1.y, d1.z, d1.w};
                            const float a2[4] = {d2.x, d2.y, d2.z, d2.w}, a3[4] = {d3.x, d3.y, d3.z, d3.w};
                            float m[4]; unsigned am4 = 0;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float c0 = a0[e] + bb[e], c1 = a1[e] + bb[e], c2 = a2[e] + bb[e], c3 = a3[e] + bb[e];
                                float mm = c0; unsigned am = 0;                  // ATen max_pool2d: (val > max) || isnan(val)
                                if (c1 > mm || c1 != c1) { mm = c1; am = 1; }
                                if (c2 > mm || c2 != c2) { mm = c2; am = 2; }
                                if (c3 > mm || c3 != c3) { mm = c3; am = 3; }
                                m[e] = mm; am4 |= am << (8 * e);
                                s1[e] += mm; s2[e] += mm * mm;
                            }
                            *reinterpret_cast<float4*>(A.out + (size_t)cq * A.out_C + A.out_off + n0) = make_float4(m[0], m[1], m[2], m[3]);
                            *reinterpret_cast<unsigned*>(A.argmax_out + (size_t)cq * A.N + n0) = am4;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                    s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                    if (psub == 0) {
                        red2[(warp * 64 + quad * 4 + e) * 2] = s1[e];
                        red2[(warp * 64 + quad * 4 + e) * 2 + 1] = s2[e];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid < 128) {
                    const int j = tid >> 1, which = tid & 1;
                    if (cb * 64 + j < A.N) {
                        double sum = 0.0;
#pragma unroll
                        for (int wq = 0; wq < 8; ++wq) sum += (double)red2[(wq * 64 + j) * 2 + which];
                        atomicAdd(A.red + ((size_t)g * A.red_C + A.out_off + cb * 64 + j) * 2 + which, sum);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");           // tile / red2 reusable
            }
        } else {
            // 64-channel column blocks: TMEM -> transposed shared tile -> lane = (pixel parity, channel quad): ReLU mask,
            // BN-backward sums (registers of the lane that owns the channel), scaled accumulate into the gradient buffer
            const int quad = lane & 15, psub = lane >> 4;
            for (int cb = 0; cb * 64 < A.Npad; ++cb) {
                {
                    const uint32_t taddr = tmem + lane_base + cb * 64 + hf * 32;
                    unsigned char* row = tb + (size_t)(q * 32 + lane) * TB_PITCH + hf * 128;
                    float v[16];
                    if (cb * 64 + hf * 32 < A.Npad) {
                        tc::tmem_ld16(taddr, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (cb * 64 + hf * 32 + 16 < A.Npad) {
                        tc::tmem_ld16(taddr + 16, v);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(row + 64 + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int n0 = cb * 64 + quad * 4;
                const bool quad_ok = n0 < A.N;
                float ca[4], cbt[4], cm[4], cs[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 t4 = quad_ok ? *reinterpret_cast<const float4*>(ntab + (n0 + e) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    ca[e] = t4.x; cbt[e] = t4.y; cm[e] = t4.z; cs[e] = t4.w;
                }
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                if (quad_ok) {
                    float4 xv[8];
                    unsigned okmask = 0u;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const long long p = p0 + warp * 16 + it * 2 + psub;
                        if (p < p_end) {
                            const size_t off = (size_t)p * A.out_C + A.out_off + n0;
                            xv[it] = __ldg(reinterpret_cast<const float4*>(A.x + off));
                            okmask |= 1u << it;
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        if (okmask & (1u << it)) {
                            const int pl = warp * 16 + it * 2 + psub;
                            const float4 d = *reinterpret_cast<const float4*>(tb + (size_t)pl * TB_PITCH + quad * 16);
                            const float4 xq = xv[it];
                            const float e0 = xq.x - cm[0], e1 = xq.y - cm[1], e2 = xq.z - cm[2], e3 = xq.w - cm[3];
                            const float y0 = fmaf(ca[0], e0, cbt[0]), y1 = fmaf(ca[1], e1, cbt[1]);
                            const float y2 = fmaf(ca[2], e2, cbt[2]), y3 = fmaf(ca[3], e3, cbt[3]);
                            const float g0 = y0 > 0.f ? d.x : 0.f;
                            const float g1 = y1 > 0.f ? d.y : 0.f;
                            const float g2 = y2 > 0.f ? d.z : 0.f;
                            const float g3 = y3 > 0.f ? d.w : 0.f;
                            if (A.a16)
                                *reinterpret_cast<uint2*>(A.a16 + (size_t)(p0 + pl) * A.N + n0) =
                                    make_uint2(tcwgrad::pack_bf16(fmaxf(y0, 0.f), fmaxf(y1, 0.f)), tcwgrad::pack_bf16(fmaxf(y2, 0.f), fmaxf(y3, 0.f)));
                            s1[0] += g0; s2[0] += g0 * (e0 * cs[0]);
                            s1[1] += g1; s2[1] += g1 * (e1 * cs[1]);
                            s1[2] += g2; s2[2] += g2 * (e2 * cs[2]);
                            s1[3] += g3; s2[3] += g3 * (e3 * cs[3]);
                            // out[p][ci] += a * g as one 16-byte L2 reduction (each element is touched once per launch)
                            tcconv::red_add_v4(A.out + (size_t)(p0 + pl) * A.out_C + A.out_off + n0, ca[0] * g0, ca[1] * g1, ca[2] * g2, ca[3] * g3);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                    s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                    if (psub == 0) {
                        red[(warp * 64 + quad * 4 + e) * 2] = s1[e];
                        red[(warp * 64 + quad * 4 + e) * 2 + 1] = s2[e];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid < 128) {
                    const int j = tid >> 1, which = tid & 1;
                    if (cb * 64 + j < A.N) {
                        double sum = 0.0;
#pragma unroll
                        for (int wq = 0; wq < 8; ++wq) sum += (double)red[(wq * 64 + j) * 2 + which];
                        atomicAdd(A.red + ((size_t)g * A.red_C + cb * 64 + j) * 2 + which, sum);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");           // tb / red reusable
            }
        }
    } else {
        // ======================================================================== MMA issuer: warp 8, convergent; one elected lane issues
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        // N > 256 is issued as two MMAs of N/2 columns (both multiples of 16)
        const int n_first = A.Npad <= 256 ? A.Npad : ((A.Npad / 2 + 15) & ~15);
        const int n_second = A.Npad - n_first;
        const uint32_t id1 = tc::instr_desc(tc::FMT_TF32, 128, n_first);
        const uint32_t id2 = n_second ? tc::instr_desc(tc::FMT_TF32, 128, n_second) : 0u;
        const uint64_t a_hi = tc::smem_desc(0, MT * 16, 128), b_hi = tc::smem_desc(0, (uint32_t)A.Npad * 16, 128);
        const uint32_t bplane = (uint32_t)A.Npad * 16;
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % NST;
            tc::mbar_wait(bars + s, (c / NST) & 1);
            tc::tc_fence_after();
            const uint32_t a_base = tc::smem_u32(a_st + s * A_STAGE), b_base = tc::smem_u32(b_st + s * bstage);
            auto mma = [&](uint32_t a_off, uint32_t b_off, uint32_t acc) {
                const uint64_t ad = a_hi | (uint64_t)((a_base + a_off) >> 4);
                tc::mma_tf32_w(tmem, ad, b_hi | (uint64_t)((b_base + b_off) >> 4), id1, acc);
                if (n_second) tc::mma_tf32_w(tmem + n_first, ad, b_hi | (uint64_t)((b_base + b_off + (uint32_t)n_first * 16u) >> 4), id2, acc);
            };
            if (A.x3) {
                const uint32_t acc = (uint32_t)(c != 0);
                mma(2u * MT * 16, 0u, acc);                    // A_lo * W_hi
                mma(0u, 2u * bplane, 1u);                      // A_hi * W_lo
                mma(0u, 0u, 1u);                               // A_hi * W_hi
            } else {
                const int nk8 = (A.K - c * 16 > 8) ? 2 : 1;
                for (int k8 = 0; k8 < nk8; ++k8) mma((uint32_t)(2 * k8) * MT * 16, (uint32_t)(2 * k8) * bplane, (uint32_t)((c | k8) != 0));
            }
            tc::tc_commit_w(bars + NST + s);
        }
        tc::tc_commit_w(bars + 2 * NST);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tc::tmem_dealloc(tmem, ncols);
    }
}

}  // namespace tcpw
}  // namespace endo
