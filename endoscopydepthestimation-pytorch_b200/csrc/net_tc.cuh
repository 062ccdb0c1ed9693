// Tensor-core (tcgen05 / TMEM) kernels of the FC-DenseNet engine for sm_100a.
//
// DenseLayer forward (reference models.py:19-28: BN -> ReLU -> conv3x3 Cin -> 12|16) as an implicit GEMM on
// the 5th-generation tensor cores:
//
//   * A operand = the BN+ReLU-transformed activation tile, produced on the fly by the staging warps while
//     they copy it from the NHWC level buffer into shared memory (no normalised copy ever exists in HBM);
//     it is stored as K-major SWIZZLE_NONE "planes": plane p holds channels 4p..4p+3 (one 16-byte chunk)
//     of every pixel of the (32+2) x (32+2) halo tile, pixels in row-major order with pitch 34, 16 bytes
//     apart.  Because rows are uniformly 16 B apart, a vertical filter tap is just a start-address offset
//     of +-34 rows in the matrix descriptor: nothing is re-staged per tap.
//   * the three horizontal taps are folded into the GEMM N dimension: N = 3 (kx) x 16 (co, 12 real) = 48,
//     so one 128 x 48 x 8 MMA consumes a 4 KB A tile for 3 taps (the shared-memory read of A, 128 B/clk,
//     is what bounds a small-N MMA).  D'[q][kx][co] = sum_ky sum_c A[q + (ky-1)*34][c] * W[co][c][ky][kx]
//     accumulates in TMEM over ky (3 descriptors) and over all input channels;
//   * the epilogue finishes the horizontal part: out[p][co] = D'[p-1][0][co] + D'[p][1][co] + D'[p+1][2][co];
//     p-1 / p+1 are the neighbouring TMEM lanes, i.e. the neighbouring threads of the warp (shuffles), with a
//     tiny shared-memory exchange at the 32-lane boundaries.  It adds the bias, writes the 12 new channels
//     in place into the level buffer and emits the per-channel sum / sum of squares for the next BatchNorms.
//
// Warp roles (288 threads): warps 0-7 stage operands (and later run the epilogue), warp 8 issues the MMAs.
// Two shared-memory stages, mbarrier full/empty pipeline, accumulators for the whole 32x32 tile (9 M-blocks
// of 128 linear pixels x 48 columns = 432 TMEM columns) stay resident in TMEM across the channel loop.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace endo {
// clock64() trace of CTA 0 of the most recent traced launch (ENDO_TC_DEBUG bit 4): see tools/trace_fwd.py
__device__ long long g_tc_trace[2048];
#define ENDO_TRACE(slot) do { if ((A.dbg & 4) && blockIdx.x == 0 && blockIdx.z == 0) g_tc_trace[(slot)] = clock64(); } while (0)

namespace tcconv {

constexpr int TH = 32, TW = 32, PITCH = 34;
constexpr int REAL_ROWS = PITCH * (TH + 2);            // 1156 halo-tile pixels
constexpr int MBLK = 9;                                // ceil(TH * PITCH / 128)
constexpr int PLANE_ROWS = 1226;                       // >= 34 + 9*128 + 34 ; 1226*16 % 128 == 32 (bank spread)
constexpr int PLANE_BYTES = PLANE_ROWS * 16;
constexpr int KCH = 16;                                // input channels per stage (tf32: 4 planes)
constexpr int A_STAGE_BYTES = (KCH / 4) * PLANE_BYTES; // 78,464
constexpr int NB = 48;                                 // MMA N = 3 kx * 16
constexpr int B_BLOCK_BYTES = 2 * NB * 16;             // one (ky, k8) weight block: 2 chunks x 48 rows x 16 B
constexpr int B_STAGE_BYTES = 3 * (KCH / 8) * B_BLOCK_BYTES;
constexpr int NUNITS = MBLK * 4;                       // (M-block, lane quadrant) epilogue units
constexpr int EDGE_FLOATS = NUNITS * 2 * 16;
constexpr int NPROD = 512;                              // 16 producer / epilogue warps + 1 MMA warp
constexpr int COEF_MAX = 1024;                          // BatchNorm coefficient table kept in shared memory (wider layers read it from L2)
constexpr int SMEM_BYTES = 2 * A_STAGE_BYTES + 2 * B_STAGE_BYTES + EDGE_FLOATS * 4 + 16 * 16 * 2 * 4 + 128 + COEF_MAX * 16;
constexpr int NTHREADS = NPROD + 32;

// kind::tf32 TRUNCATES the low 13 mantissa bits of its fp32 operands (tests/test_gpu_tc_probe.py).  Truncation is biased
// (every magnitude shrinks by ~2^-11 on average) and the bias compounds through the 57 stacked layers, so every operand
// is rounded to nearest tf32 while it is staged: the rounding error is then zero-mean and averages out over the GEMM K.
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// 16-byte read-only load that asks the L2 to fetch the whole 256-byte-aligned block around it from DRAM.  The forward producers
// walk the channels of a pixel chunk by chunk (32 or 64 bytes of a pixel per pipeline step, pixels 4*C bytes apart): without
// the hint every step is one scattered 32-byte DRAM sector per pixel; with it the DRAM access is 256 contiguous bytes and the
// following chunks of the pixel hit the L2.  (Explicit prefetch.global.L2 instructions several chunks ahead were measured
// SLOWER, +5-10 % per kernel: they are cache-control operations competing for the same issue / LSU slots.)
__device__ __forceinline__ float4 ldg_l2_256(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// v += d as ONE 16-byte reduction performed by the L2 (REDG.E.ADD.F32x4): the read-modify-write of a gradient tile no longer
// round-trips through the SM (no load to wait for, no registers held across the wait)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// hi part of the 3xTF32 split v = hi + lo (lo = v - hi is exact in fp32 and fits the 11 bits tf32 keeps up to 2^-23 |v|)
__device__ __forceinline__ float tf32_hi(float v) { return tf32_rn(v); }
__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Weight images: the exact shared-memory layout of one pipeline stage, built once per pass for ALL layers by
// pack_w_fwd_all_kernel / pack_w_dgrad_all_kernel (the 1x1 fall-back path packs per layer) so
// that the producers copy them with coalesced 128-bit loads (staging OIHW weights with scalar, serialised loads cost
// more than the activation stream in the first version of these kernels).
//   forward : chunk c (16 input channels) -> [blk = ky*2 + k8][kc][n = kx*16 + co][4]           2304 floats
//   dgrad   : chunk c (64 input channels) -> [blk = tap*2 + k8][kc][n = ci - 64c][4], taps flipped 9216 floats

// 3xTF32 (error-compensated, fp32-grade) variants: a chunk holds 8 input channels.  x = hi + lo with hi = x rounded to tf32;
// x*w = hi*hi + (lo*w + x*lo_w) + O(2^-24): block (ky, 0) carries the tf32 hi weights for ONE kind::tf32 MMA (K = 8), block
// (ky, 1) the bf16 pair [w (8 channels) ; w - hi (8 channels)] for ONE kind::f16 MMA of K = 16 that adds BOTH cross terms
// against the activation planes [lo ; x] (the cross terms are 2^-12 of the product, so 8-bit operands keep them to 2^-21).
__global__ void __launch_bounds__(256)
pack_w_1x1_x3_kernel(const float* __restrict__ w, int K, int Ntot, int co0, float* __restrict__ out) {
    pdl_enter();
    const int c = blockIdx.x;
    for (int d = threadIdx.x; d < 768; d += 256) {
        const int part = d / 384, r = d - part * 384, kc = r / 192, r2 = r - kc * 192, n = r2 >> 2, e = r2 & 3;
        const int cin = c * 8 + kc * 4 + e, co = co0 + n;
        float v = 0.f;
        if (co < Ntot && cin < K) v = __ldg(w + (size_t)co * K + cin);
        if (part == 0) { out[(size_t)c * 2304 + d] = tf32_hi(v); continue; }
        const int c0 = c * 8 + 2 * e;
        float w0 = 0.f, w1 = 0.f;
        if (co < Ntot && c0 < K) w0 = __ldg(w + (size_t)co * K + c0);
        if (co < Ntot && c0 + 1 < K) w1 = __ldg(w + (size_t)co * K + c0 + 1);
        if (kc) { w0 -= tf32_hi(w0); w1 -= tf32_hi(w1); }
        out[(size_t)c * 2304 + d] = __uint_as_float(bf16x2_rn(w0, w1));
    }
}

// bf16x3 (error-compensated bf16: x = b1 + b2, two bf16 terms = 16 significant bits, three kind::f16 MMAs of K = 16 per
// product -- half the MMA count of 3xTF32): a chunk holds 16 input channels, block (ky, term) = [kc (8 channels)][n][8 bf16].
// returns the packed first terms; (rlo, rhi) receive the exact remainders
__device__ __forceinline__ uint32_t bf16_split(float lo, float hi, float& rlo, float& rhi) {
    const uint32_t p = bf16x2_rn(lo, hi);
    rlo = lo - __uint_as_float(p << 16);
    rhi = hi - __uint_as_float(p & 0xFFFF0000u);
    return p;
}
__global__ void __launch_bounds__(256)
pack_w_1x1_b3_kernel(const float* __restrict__ w, int K, int Ntot, int co0, uint32_t* __restrict__ out) {
    pdl_enter();
    const int c = blockIdx.x;
    for (int d = threadIdx.x; d < 768; d += 256) {
        const int term = d / 384, r = d - term * 384, kc = r / 192, r2 = r - kc * 192, n = r2 >> 2, e = r2 & 3;
        const int cin = c * 16 + kc * 8 + e * 2, co = co0 + n;
        float v0 = 0.f, v1 = 0.f;
        if (co < Ntot && cin < K) v0 = __ldg(w + (size_t)co * K + cin);
        if (co < Ntot && cin + 1 < K) v1 = __ldg(w + (size_t)co * K + cin + 1);
        float r0, r1;
        const uint32_t b1 = bf16_split(v0, v1, r0, r1);
        out[(size_t)c * 2304 + d] = term ? bf16x2_rn(r0, r1) : b1;
    }
}

// 1x1 convolution (TransitionDown): only blocks k8 = 0, 1 of the stage image are used, n = output channel co0 + n
__global__ void __launch_bounds__(256)
pack_w_1x1_kernel(const float* __restrict__ w, int K, int Ntot, int co0, float* __restrict__ out) {
    pdl_enter();
    const int c = blockIdx.x;
    for (int d = threadIdx.x; d < 2304; d += 256) {
        const int blk = d / 384, r = d - blk * 384, kc = r / 192, r2 = r - kc * 192, n = r2 >> 2, e = r2 & 3;
        const int cin = c * 16 + blk * 8 + kc * 4 + e, co = co0 + n;
        float v = 0.f;
        if (blk < 2 && co < Ntot && cin < K) v = __ldg(w + (size_t)co * K + cin);
        out[(size_t)c * 2304 + d] = tf32_rn(v);
    }
}

// 2x2 max-pool of the TransitionDown conv output (tensor-core path): pooled value -> next level buffer, argmax byte,
// per-channel statistics.  HBM-bound elementwise pass (reads 16 B/px/channel-quad, writes 4).  A warp walks coarse
// pixels, lane = channel quad (q = lane, lane + 32, ...), so a pixel is read as one contiguous run and the statistics
// of a channel are accumulated by ONE lane in a fixed order (bit-reproducible forward, like the FFMA epilogue).
constexpr int POOL_MAXQ = 6;                               // channel quads per lane: Cs <= 768
__global__ void __launch_bounds__(256)
td_pool_kernel(const float* __restrict__ tmp, float* __restrict__ out, unsigned char* __restrict__ argmax, double* __restrict__ stats,
               int B, int h, int w, int Cs, int out_C, int out_off, int G, int stats_C) {
    pdl_enter();
    extern __shared__ float sm[];                          // [8 warps][Cs][2]
    const int nq = Cs >> 2, oh = h >> 1, ow = w >> 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long per_group = (long long)(B / G) * oh * ow;
    const int g = blockIdx.y;
    float s1[POOL_MAXQ][4], s2[POOL_MAXQ][4];
#pragma unroll
    for (int qi = 0; qi < POOL_MAXQ; ++qi)
#pragma unroll
        for (int e = 0; e < 4; ++e) { s1[qi][e] = 0.f; s2[qi][e] = 0.f; }
    for (long long i = (long long)blockIdx.x * 8 + warp; i < per_group; i += (long long)gridDim.x * 8) {
        const long long pp = (long long)g * per_group + i;               // coarse pixel (b, y2, x2) flattened
        const int x2 = (int)(pp % ow), y2 = (int)((pp / ow) % oh), b = (int)(pp / ((long long)ow * oh));
        const float* p00 = tmp + (((size_t)b * h + 2 * y2) * w + 2 * x2) * Cs;
#pragma unroll
        for (int qi = 0; qi < POOL_MAXQ; ++qi) {
            const int q = lane + 32 * qi;
            if (q < nq) {
                const float* pq = p00 + q * 4;
                const float4 v00 = __ldg(reinterpret_cast<const float4*>(pq)), v01 = __ldg(reinterpret_cast<const float4*>(pq + Cs));
                const float4 v10 = __ldg(reinterpret_cast<const float4*>(pq + (size_t)w * Cs));
                const float4 v11 = __ldg(reinterpret_cast<const float4*>(pq + (size_t)w * Cs + Cs));
                const float a0[4] = {v00.x, v00.y, v00.z, v00.w}, a1[4] = {v01.x, v01.y, v01.z, v01.w};
                const float a2[4] = {v10.x, v10.y, v10.z, v10.w}, a3[4] = {v11.x, v11.y, v11.z, v11.w};
                float m[4]; unsigned am4 = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float mm = a0[e]; unsigned am = 0;                  // ATen max_pool2d: (val > max) || isnan(val)
                    if (a1[e] > mm || a1[e] != a1[e]) { mm = a1[e]; am = 1; }
                    if (a2[e] > mm || a2[e] != a2[e]) { mm = a2[e]; am = 2; }
                    if (a3[e] > mm || a3[e] != a3[e]) { mm = a3[e]; am = 3; }
                    m[e] = mm; am4 |= am << (8 * e);
                    s1[qi][e] += mm; s2[qi][e] += mm * mm;
                }
                *reinterpret_cast<float4*>(out + (size_t)pp * out_C + out_off + q * 4) = make_float4(m[0], m[1], m[2], m[3]);
                *reinterpret_cast<unsigned*>(argmax + (size_t)pp * Cs + q * 4) = am4;
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < POOL_MAXQ; ++qi) {
        const int q = lane + 32 * qi;
        if (q < nq) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sm[((size_t)warp * Cs + q * 4 + e) * 2] = s1[qi][e];
                sm[((size_t)warp * Cs + q * 4 + e) * 2 + 1] = s2[qi][e];
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cs * 2; i += 256) {
        double sum = 0.0;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) sum += (double)sm[(size_t)wq * Cs * 2 + i];
        atomicAdd(stats + ((size_t)g * stats_C + out_off + (i >> 1)) * 2 + (i & 1), sum);
    }
}

// All DenseLayer weight images of one forward (or backward) in ONE launch: the weights do not change inside a step, and
// 44 + 44 tiny per-layer pack launches sat on the critical path of the launch-serialised layer chain.
struct PackEntry { long long w; int K, N, chunk0; long long out; };      // w: float offset in params; out: byte offset in the wpack region
struct PackTable { int n, total_chunks, mode; PackEntry e[112]; };        // mode: forward 0 = tf32, 1 = 3xTF32, 2 = bf16x3; dgrad: unused
__global__ void __launch_bounds__(256)
pack_w_fwd_all_kernel(const float* __restrict__ params, unsigned char* __restrict__ wpack, const PackTable T) {
    pdl_enter();
    int li = 0;
    while (li + 1 < T.n && (int)blockIdx.x >= T.e[li + 1].chunk0) ++li;
    const PackEntry E = T.e[li];
    const int c = blockIdx.x - E.chunk0, K = E.K, N = E.N;
    const float* w = params + E.w;
    float* out = reinterpret_cast<float*>(wpack + E.out);
    for (int d = threadIdx.x; d < 2304; d += 256) {
        const int blk = d / 384, r = d - blk * 384, kc = r / 192, r2 = r - kc * 192, n = r2 >> 2, e = r2 & 3;
        const int ky = blk >> 1, part = blk & 1, co = n & 15, kx = n >> 4;
        float res;
        if (T.mode == 0) {                                                 // plain tf32: 16 input channels per chunk, block (ky, k8)
            const int cin = c * 16 + part * 8 + kc * 4 + e;
            float v = 0.f;
            if (co < N && cin < K) v = __ldg(w + (((size_t)co * K + cin) * 3 + ky) * 3 + kx);
            res = tf32_rn(v);
        } else if (T.mode == 1 && part == 0) {                             // 3xTF32: tf32 hi block of the 8-channel chunk
            const int cin = c * 8 + kc * 4 + e;
            float v = 0.f;
            if (co < N && cin < K) v = __ldg(w + (((size_t)co * K + cin) * 3 + ky) * 3 + kx);
            res = tf32_hi(v);
        } else {                                                           // bf16 blocks of 3xTF32 (cross terms) and of bf16x3
            const int c0 = (T.mode == 1) ? c * 8 + 2 * e : c * 16 + kc * 8 + 2 * e;
            float w0 = 0.f, w1 = 0.f;
            if (co < N && c0 < K) w0 = __ldg(w + (((size_t)co * K + c0) * 3 + ky) * 3 + kx);
            if (co < N && c0 + 1 < K) w1 = __ldg(w + (((size_t)co * K + c0 + 1) * 3 + ky) * 3 + kx);
            if (T.mode == 1) {                                             // kc = 0: w, kc = 1: w - hi
                if (kc) { w0 -= tf32_hi(w0); w1 -= tf32_hi(w1); }
                res = __uint_as_float(bf16x2_rn(w0, w1));
            } else {                                                       // part = 0: first bf16 terms, part = 1: remainders
                const uint32_t b1 = bf16x2_rn(w0, w1);
                const float r0 = w0 - __uint_as_float(b1 << 16), r1 = w1 - __uint_as_float(b1 & 0xFFFF0000u);
                res = __uint_as_float(part ? bf16x2_rn(r0, r1) : b1);
            }
        }
        out[(size_t)c * 2304 + d] = res;
    }
}
__global__ void __launch_bounds__(256)
pack_w_dgrad_all_kernel(const float* __restrict__ params, unsigned char* __restrict__ wpack, const PackTable T) {
    pdl_enter();
    int li = 0;
    while (li + 1 < T.n && (int)blockIdx.x >= T.e[li + 1].chunk0) ++li;
    const PackEntry E = T.e[li];
    const int c = blockIdx.x - E.chunk0, Cin = E.K, Cout = E.N;
    const float* w = params + E.w;
    float* out = reinterpret_cast<float*>(wpack + E.out);
    for (int d = threadIdx.x; d < 9216; d += 256) {                        // chunk c (64 input channels) -> [blk = tap*2 + k8][kc][n = ci - 64c][4], taps flipped
        const int blk = d >> 9, r = d & 511, kc = r >> 8, n = (r & 255) >> 2, e = r & 3;
        const int tap = blk >> 1, k8 = blk & 1, co = k8 * 8 + kc * 4 + e, ci = c * 64 + n;
        float v = 0.f;
        if (co < Cout && ci < Cin) v = __ldg(w + ((size_t)co * Cin + ci) * 9 + (8 - tap));
        out[(size_t)c * 9216 + d] = tf32_rn(v);
    }
}

struct FwdArgs {
    const float* in; const float* coef; const float* w; const float* bias; float* out; double* stats;
    int in_C, in_off, K, out_C, out_off, N, H, W, B, G, stats_C;
    int up;      // 0: operand = relu(bn(x)) of the same-resolution buffer (DenseLayer); 1: operand = x of the half-resolution
                 //    buffer, nearest-upsampled x2, no BatchNorm (TransitionUp, models.py:73-74)
    int dbg;     // performance experiments only (ENDO_TC_DEBUG): 1 = skip the MMAs, 2 = skip the activation loads
    const float* wpack;   // weight image built by pack_w_fwd_all_kernel (2304 floats per chunk)
    int one;              // 1: 1x1 convolution, N <= 48 plain output channels, epilogue = + bias, store (no taps, no statistics)
    float* partial; int ksplit; long long pixels;   // split-K (not with `one`): [ksplit][B*H*W][16] raw partial sums, else nullptr
    int x3;               // 1: 3xTF32 -- a stage holds 8 channels as planes (tf32 hi0, hi1; bf16 lo; bf16 x): D += [lo;x]*[w;wlo] + hi*whi
                          // 2: bf16x3 -- a stage holds 16 channels as bf16 planes (b1: ch 0-7, 8-15; b2: ch 0-7, 8-15), kind::f16
};

__global__ void __launch_bounds__(NTHREADS, 1)
dense_fwd_tf32_kernel(const FwdArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* a_st0 = smem;
    unsigned char* b_st0 = smem + 2 * A_STAGE_BYTES;
    float* edge = reinterpret_cast<float*>(smem + 2 * A_STAGE_BYTES + 2 * B_STAGE_BYTES);   // [NUNITS][2][16]
    float* red = edge + EDGE_FLOATS;                                                        // [16][16][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16 * 16 * 2);                         // full[2], empty[2], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    float* coef_s = reinterpret_cast<float*>(smem + 2 * A_STAGE_BYTES + 2 * B_STAGE_BYTES + EDGE_FLOATS * 4 + 16 * 16 * 2 * 4 + 128);
    __shared__ float s_bias[48];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (A.W + TW - 1) / TW;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    const int b = blockIdx.z;
    const int g = b / (A.B / A.G);
    const int kch = (A.x3 == 1) ? 8 : KCH;                            // input channels per pipeline stage
    const int nchunks = (A.K + kch - 1) / kch;
    // split-K (low-resolution levels: few tiles, long channel loops): blockIdx.y owns a slice of the channel chunks and
    // stores raw partial sums; splitk_finish_kernel adds the slices in a fixed order, then bias + statistics
    const int per_slice = A.partial ? (nchunks + A.ksplit - 1) / A.ksplit : nchunks;
    const int c_begin = A.partial ? (int)blockIdx.y * per_slice : 0;
    const int c_end = min(nchunks, c_begin + per_slice);

    pdl_trigger();
    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        tc::mbar_init(bars + 0, NPROD); tc::mbar_init(bars + 1, NPROD);
        tc::mbar_init(bars + 2, 1);   tc::mbar_init(bars + 3, 1);
        tc::mbar_init(bars + 4, 1);
        tc::fence_mbar_init();
    }
    pdl_wait();                        // everything above is on-chip; global memory only from here on
    if (tid == 0) ENDO_TRACE(1);
    // (a, beta, mean, invstd) of every input channel of this statistic group: one L2 read per CTA instead of four dependent
    // global loads per thread and channel chunk (r2 ncu: their latency sat on the producers' critical path)
    const bool coef_smem = !A.up && A.K <= COEF_MAX;
    if (coef_smem) {
        const int gq = (int)blockIdx.z / (A.B / A.G);
        for (int i = tid; i < A.K; i += NTHREADS)
            *reinterpret_cast<float4*>(coef_s + i * 4) = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)gq * A.K + i) * 4));
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 16) {
        // bias is needed only in the epilogue, but a global load issued there would pay the loaded-DRAM latency
        // (~4 us while every SM streams operands): fetch it now
        if (tid < 48) s_bias[tid] = (tid < A.N) ? __ldg(A.bias + tid) : 0.f;     // visible after the named barrier below
        // ======================================================================== producers
        if (A.x3 == 1) {
            // ---- 3xTF32 producer: a stage = 8 channels as planes (hi0, hi1, lo0, lo1).  An item = (pixel, 4-channel group):
            // ONE 16-byte load, BN+ReLU, exact split, TWO 16-byte shared stores.  Software-pipelined: the loads of chunk c+1
            // are in flight (registers) while chunk c is transformed and stored, so the memory system never idles between
            // stages; pixel offsets / validity are loop-invariant and computed once.
            constexpr int NIT = 5;                                    // ceil(2 * 1156 / 512)
            const int cg = tid & 1;
            const int sh = A.up ? 1 : 0;
            const int sW = A.W >> sh;
            const float* in_img = A.in + (size_t)b * (A.H >> sh) * sW * A.in_C + A.in_off + cg * 4;
            uint32_t poff[NIT];
            unsigned pixok = 0u, rowok = 0u;
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                const int px = (tid >> 1) + 256 * j;
                const int r = px / PITCH, cc = px - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                poff[j] = 0u;
                if (px < REAL_ROWS) {
                    rowok |= 1u << j;
                    if (y >= 0 && y < A.H && x >= 0 && x < A.W && !(A.dbg & 2)) {
                        pixok |= 1u << j;
                        poff[j] = (uint32_t)(((y >> sh) * sW + (x >> sh)) * A.in_C);
                    }
                }
            }
            float4 cur[NIT], nxt[NIT];
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                cur[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((pixok & (1u << j)) && c_begin * 8 + cg * 4 < A.K)
                    cur[j] = ldg_l2_256(in_img + c_begin * 8 + poff[j]);
            }
            for (int c = c_begin; c < c_end; ++c) {
                const int it = c - c_begin;
                const int s = it & 1;
                const int ch = c * 8 + cg * 4;
                const bool ch_ok = ch < A.K;
                const bool nxt_ok = ch + 8 < A.K && c + 1 < c_end;
#pragma unroll
                for (int j = 0; j < NIT; ++j) {                        // prefetch chunk c + 1
                    nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((pixok & (1u << j)) && nxt_ok) nxt[j] = ldg_l2_256(in_img + (c + 1) * 8 + poff[j]);
                }

                float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0, k3 = k0;   // (a, beta, mean, invstd) x 4 channels
                if (ch_ok && !A.up) {
                    if (coef_smem) {
                        const float* cf = coef_s + ch * 4;
                        k0 = *reinterpret_cast<const float4*>(cf); k1 = *reinterpret_cast<const float4*>(cf + 4);
                        k2 = *reinterpret_cast<const float4*>(cf + 8); k3 = *reinterpret_cast<const float4*>(cf + 12);
                    } else {
                        const float* cf = A.coef + ((size_t)g * A.K + ch) * 4;
                        k0 = __ldg(reinterpret_cast<const float4*>(cf)); k1 = __ldg(reinterpret_cast<const float4*>(cf + 4));
                        k2 = __ldg(reinterpret_cast<const float4*>(cf + 8)); k3 = __ldg(reinterpret_cast<const float4*>(cf + 12));
                    }
                }
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 0);
                if (it >= 2) tc::mbar_wait(bars + 2 + s, ((it >> 1) - 1) & 1);
                // weights of this chunk: the ready-made 9,216-byte stage image travels by TMA bulk copy straight into the stage
                // (one thread, no registers), completing on the stage's "full" barrier as a transaction count
                if (tid == 0) {
                    tc::mbar_expect_tx(bars + s, (uint32_t)B_STAGE_BYTES);
                    tc::bulk_g2s(b_st0 + s * B_STAGE_BYTES, A.wpack + (size_t)c * 2304, (uint32_t)B_STAGE_BYTES, bars + s);
                }
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 1);
                unsigned char* a_hi_s = a_st0 + s * A_STAGE_BYTES + cg * PLANE_BYTES;
#pragma unroll
                for (int j = 0; j < NIT; ++j) {
                    if (rowok & (1u << j)) {
                        const int px = (tid >> 1) + 256 * j;
                        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);
                        uint2 lo = make_uint2(0u, 0u), xb = lo;
                        if ((pixok & (1u << j)) && ch_ok) {
                            float4 v = cur[j];
                            if (!A.up) {
                                v.x = fmaxf(fmaf(k0.x, v.x - k0.z, k0.y), 0.f); v.y = fmaxf(fmaf(k1.x, v.y - k1.z, k1.y), 0.f);
                                v.z = fmaxf(fmaf(k2.x, v.z - k2.z, k2.y), 0.f); v.w = fmaxf(fmaf(k3.x, v.w - k3.z, k3.y), 0.f);
                            }
                            hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                            lo = make_uint2(bf16x2_rn(v.x - hi.x, v.y - hi.y), bf16x2_rn(v.z - hi.z, v.w - hi.w));
                            xb = make_uint2(bf16x2_rn(v.x, v.y), bf16x2_rn(v.z, v.w));
                        }
                        // plane cg: tf32 hi (4 channels); plane 2: bf16 lo, plane 3: bf16 x (8 channels per 16-byte row, this
                        // thread's 4 channels = bytes 8 cg .. 8 cg + 7)
                        *reinterpret_cast<float4*>(a_hi_s + (size_t)px * 16) = hi;
                        *reinterpret_cast<uint2*>(a_st0 + s * A_STAGE_BYTES + 2 * PLANE_BYTES + (size_t)px * 16 + cg * 8) = lo;
                        *reinterpret_cast<uint2*>(a_st0 + s * A_STAGE_BYTES + 3 * PLANE_BYTES + (size_t)px * 16 + cg * 8) = xb;
                    }
                }
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 2);
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 3);
                tc::fence_proxy_async();
                tc::mbar_arrive(bars + s);
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 4);
#pragma unroll
                for (int j = 0; j < NIT; ++j) cur[j] = nxt[j];
            }
        }
        const int grp = tid & 3;                                  // this thread always stages the same plane (4-channel group)
        const int cgrp = grp;
        for (int c = c_begin; c < (A.x3 == 1 ? c_begin : c_end); ++c) {
            const int it = c - c_begin;
            const int s = it & 1;
            if (tid == 0) ENDO_TRACE(16 + c * 8 + 0);
            unsigned char* a_s = a_st0 + s * A_STAGE_BYTES + grp * PLANE_BYTES;
            const int ch = c * kch + cgrp * 4;
            float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0, k3 = k0;   // (a, beta, mean, invstd) x 4 channels
            const bool ch_ok = ch < A.K;
            if (ch_ok && !A.up) {
                if (coef_smem) {
                    const float* cf = coef_s + ch * 4;
                    k0 = *reinterpret_cast<const float4*>(cf); k1 = *reinterpret_cast<const float4*>(cf + 4);
                    k2 = *reinterpret_cast<const float4*>(cf + 8); k3 = *reinterpret_cast<const float4*>(cf + 12);
                } else {
                    const float* cf = A.coef + ((size_t)g * A.K + ch) * 4;
                    k0 = __ldg(reinterpret_cast<const float4*>(cf)); k1 = __ldg(reinterpret_cast<const float4*>(cf + 4));
                    k2 = __ldg(reinterpret_cast<const float4*>(cf + 8)); k3 = __ldg(reinterpret_cast<const float4*>(cf + 12));
                }
            }
            const int sh = A.up ? 1 : 0;                              // source = (y >> sh, x >> sh) of a (H >> sh) x (W >> sh) buffer
            const int sW = A.W >> sh;
            const float* in_b = A.in + (size_t)b * (A.H >> sh) * sW * A.in_C + A.in_off + ch;
            // 10 pixels per thread (512 producer threads, 4 per pixel): ONE batch of independent 16-byte loads, i.e. one
            // memory round trip per stage with 80 KB per SM in flight (the kernel is bound by bytes in flight)
            {
                float4 q[10];
                unsigned okmask = 0u;
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const int px = (tid >> 2) + 128 * j;
                    const int r = px / PITCH, cc = px - r * PITCH;
                    const int y = y0 + r - 1, x = x0 + cc - 1;
                    const bool ok = ch_ok && (px < REAL_ROWS) && y >= 0 && y < A.H && x >= 0 && x < A.W && !(A.dbg & 2);
                    q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) {
                        q[j] = ldg_l2_256(in_b + ((size_t)(y >> sh) * sW + (x >> sh)) * A.in_C);
                        okmask |= 1u << j;
                    }
                }
                // the loads above are in flight while this thread waits for the MMAs that still read the stage
                if (it >= 2) tc::mbar_wait(bars + 2 + s, ((it >> 1) - 1) & 1);
                if (tid == 0) {                                    // weights of this chunk: TMA bulk copy of the 9,216-byte stage image
                    tc::mbar_expect_tx(bars + s, (uint32_t)B_STAGE_BYTES);
                    tc::bulk_g2s(b_st0 + s * B_STAGE_BYTES, A.wpack + (size_t)c * 2304, (uint32_t)B_STAGE_BYTES, bars + s);
                }
                if (tid == 0) ENDO_TRACE(16 + c * 8 + 1);
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const int px = (tid >> 2) + 128 * j;
                    if (px < REAL_ROWS) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (okmask & (1u << j)) {
                            if (A.up) v = q[j];
                            else {
                                v.x = fmaxf(fmaf(k0.x, q[j].x - k0.z, k0.y), 0.f); v.y = fmaxf(fmaf(k1.x, q[j].y - k1.z, k1.y), 0.f);
                                v.z = fmaxf(fmaf(k2.x, q[j].z - k2.z, k2.y), 0.f); v.w = fmaxf(fmaf(k3.x, q[j].w - k3.z, k3.y), 0.f);
                            }
                            if (A.x3 == 0) v = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
                        }
                        if (A.x3 >= 2) {
                            // bf16x3 / bf16: this thread's 4 channels are 8 bytes of the 16-byte (8-channel) row of plane grp / 2 (first
                            // terms) and of plane 2 + grp / 2 (second terms = bf16 of the exact remainders)
                            float r0, r1, r2, r3;
                            const uint32_t p0 = bf16_split(v.x, v.y, r0, r1), p1 = bf16_split(v.z, v.w, r2, r3);
                            unsigned char* row = a_st0 + s * A_STAGE_BYTES + (grp >> 1) * PLANE_BYTES + (size_t)px * 16 + (grp & 1) * 8;
                            *reinterpret_cast<uint2*>(row) = make_uint2(p0, p1);
                            *reinterpret_cast<uint2*>(row + 2 * PLANE_BYTES) = make_uint2(bf16x2_rn(r0, r1), bf16x2_rn(r2, r3));
                        } else
                        *reinterpret_cast<float4*>(a_s + (size_t)px * 16) = v;
                    }
                }
            }
            if (tid == 0) ENDO_TRACE(16 + c * 8 + 2);
            if (tid == 0) ENDO_TRACE(16 + c * 8 + 3);
            tc::fence_proxy_async();
            tc::mbar_arrive(bars + s);
            if (tid == 0) ENDO_TRACE(16 + c * 8 + 4);
        }
        // ======================================================================== epilogue
        if (tid == 0) ENDO_TRACE(2);
        tc::mbar_wait(bars + 4, 0);
        tc::tc_fence_after();
        if (tid == 0) ENDO_TRACE(3);
        const int q = warp & 3, part = warp >> 2;             // TMEM lane quadrant, and which M-blocks this warp drains
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        if (A.one) {
            asm volatile("bar.sync 1, 512;" ::: "memory");       // s_bias visible
            for (int mb = part; mb < MBLK; mb += 4) {
                const int L = PITCH + mb * 128 + q * 32 + lane;
                const int r = L / PITCH, cc = L - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                const bool ok = (r <= TH) && (cc >= 1) && (cc <= TW) && (y < A.H) && (x < A.W);
                float* op = A.out + ((size_t)(b * A.H + y) * A.W + x) * A.out_C + A.out_off;
#pragma unroll 1
                for (int c16 = 0; c16 < 3; ++c16) {
                    float v[16];
                    tc::tmem_ld16(tmem + lane_base + mb * NB + c16 * 16, v);
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            if (c16 * 16 + j < A.N)
                                *reinterpret_cast<float4*>(op + c16 * 16 + j) =
                                    make_float4(v[j] + s_bias[c16 * 16 + j], v[j + 1] + s_bias[c16 * 16 + j + 1],
                                                v[j + 2] + s_bias[c16 * 16 + j + 2], v[j + 3] + s_bias[c16 * 16 + j + 3]);
                        }
                    }
                }
            }
        } else {
        // pass 1: publish the values the neighbouring 32-lane units need
        for (int mb = part; mb < MBLK; mb += 4) {
            const int u = mb * 4 + q;
            uint32_t r0[16], r2[16];
            tc::tmem_ld16_issue(tmem + lane_base + mb * NB + 0, r0);
            tc::tmem_ld16_issue(tmem + lane_base + mb * NB + 32, r2);
            tc::tmem_ld_wait();
            if (lane == 31) {
#pragma unroll
                for (int j = 0; j < 16; ++j) edge[(u * 2 + 1) * 16 + j] = __uint_as_float(r0[j]);    // kx = 0 part of my last pixel
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) edge[(u * 2 + 0) * 16 + j] = __uint_as_float(r2[j]);    // kx = 2 part of my first pixel
            }
        }
        if (tid == 0) ENDO_TRACE(6);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (tid == 0) ENDO_TRACE(7);
        float s1[16], s2[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        for (int mb = part; mb < MBLK; mb += 4) {
            const int u = mb * 4 + q;
            uint32_t v0[16], v1[16], v2[16];
            tc::tmem_ld16_issue(tmem + lane_base + mb * NB + 0, v0);
            tc::tmem_ld16_issue(tmem + lane_base + mb * NB + 16, v1);
            tc::tmem_ld16_issue(tmem + lane_base + mb * NB + 32, v2);
            tc::tmem_ld_wait();
            if (tid == 0) ENDO_TRACE(9 + (mb >> 2) * 2);
            const int L = PITCH + mb * 128 + q * 32 + lane;           // linear index in the halo tile
            const int r = L / PITCH, cc = L - r * PITCH;
            const int y = y0 + r - 1, x = x0 + cc - 1;
            const bool ok = (r <= TH) && (cc >= 1) && (cc <= TW) && (y < A.H) && (x < A.W);
            float o[16];
            // values of the neighbouring 32-lane units: uniform (broadcast) 16-byte loads + selects.  (Per-channel `if (lane == 0)`
            // loads were 32 divergent regions per M-block: ~4k of the ~5.5k cycles an M-block took here, clock64 trace.)
            const float* eL = edge + ((u > 0 ? u - 1 : 0) * 2 + 1) * 16;
            const float* eR = edge + ((u < NUNITS - 1 ? u + 1 : u) * 2 + 0) * 16;
            const float keepL = (u > 0) ? 1.f : 0.f, keepR = (u < NUNITS - 1) ? 1.f : 0.f;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(eL + j4), r4 = *reinterpret_cast<const float4*>(eR + j4);
                const float lq[4] = {l4.x * keepL, l4.y * keepL, l4.z * keepL, l4.w * keepL};
                const float rq[4] = {r4.x * keepR, r4.y * keepR, r4.z * keepR, r4.w * keepR};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = j4 + e;
                    float left = __uint_as_float(__shfl_up_sync(0xffffffffu, v0[j], 1));
                    float right = __uint_as_float(__shfl_down_sync(0xffffffffu, v2[j], 1));
                    left = (lane == 0) ? lq[e] : left;
                    right = (lane == 31) ? rq[e] : right;
                    o[j] = (left + __uint_as_float(v1[j])) + right + (A.partial ? 0.f : s_bias[j]);
                }
            }
            if (ok && A.partial) {
                float* pp = A.partial + ((size_t)blockIdx.y * A.pixels + ((size_t)(b * A.H + y) * A.W + x)) * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(pp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            } else if (ok) {
                float* op = A.out + ((size_t)(b * A.H + y) * A.W + x) * A.out_C + A.out_off;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    if (j < A.N) {
                        if (!(A.dbg & 16)) *reinterpret_cast<float4*>(op + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { s1[j + e] += o[j + e]; s2[j + e] += o[j + e] * o[j + e]; }
                    }
                }
            }
            if (tid == 0) ENDO_TRACE(10 + (mb >> 2) * 2);
        }
        if (tid == 0) ENDO_TRACE(8);
        // per-channel statistics: warp tree -> shared -> one fp64 atomic per channel per CTA
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float a = s1[j], c2 = s2[j];
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o2);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o2);
            }
            if (lane == 0) { red[(warp * 16 + j) * 2] = a; red[(warp * 16 + j) * 2 + 1] = c2; }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (tid < 2 * A.N && !A.partial) {
            const int j = tid >> 1, which = tid & 1;
            double sum = 0.0;
#pragma unroll
            for (int wq = 0; wq < 16; ++wq) sum += (double)red[(wq * 16 + j) * 2 + which];
            atomicAdd(A.stats + ((size_t)g * A.stats_C + A.out_off + j) * 2 + which, sum);
        }
        }   // !A.one
    } else {
        // ======================================================================== MMA issuer: warp 16, convergent; one elected lane issues
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t idesc = tc::instr_desc(tc::FMT_TF32, 128, NB);
        for (int c = c_begin; c < c_end; ++c) {
            const int it = c - c_begin;
            const int s = it & 1;
            if (lane == 0) ENDO_TRACE(16 + c * 8 + 5);
            tc::mbar_wait(bars + s, (it >> 1) & 1);
            tc::tc_fence_after();
            if (lane == 0) ENDO_TRACE(16 + c * 8 + 6);
            const uint32_t a_base = tc::smem_u32(a_st0 + s * A_STAGE_BYTES);
            const uint32_t b_base = tc::smem_u32(b_st0 + s * B_STAGE_BYTES);
            const int nk8 = (A.x3 || A.K - c * KCH > 8) ? 2 : 1;   // skip an all-zero K half on the last chunk
            const uint32_t idesc_b = tc::instr_desc(tc::FMT_BF16, 128, NB);
            // A tcgen05.mma that accumulates into the SAME TMEM tile as its predecessor waits ~266 cycles for it
            // (measured, tests/test_gpu_tc_probe.py::test_mma_cost_by_operand_layout), whatever its size.  Consecutive
            // MMAs therefore target different M-blocks: 9 independent accumulator chains keep the pipe busy.
            const uint64_t a_hi = tc::smem_desc(0, PLANE_BYTES, 128), b_hi = tc::smem_desc(0, NB * 16, 128);
            if (A.x3) {
                // 3xTF32: planes (0,1) = A_hi, planes (2,3) = A_lo; weight block (ky, 0) = W_hi, (ky, 1) = W_lo.  The small
                // cross terms are accumulated first, the hi*hi term last.
                const int nky = A.one ? 1 : 3;
#pragma unroll 1
                for (int ky = 0; ky < nky; ++ky) {
                    const uint32_t row0 = (uint32_t)(A.one ? PITCH : PITCH + (ky - 1) * PITCH) * 16u;
                    const uint64_t ahi = a_hi | (uint64_t)((a_base + row0) >> 4);
                    const uint64_t alo = a_hi | (uint64_t)((a_base + 2u * PLANE_BYTES + row0) >> 4);
                    const uint64_t bhi = b_hi | (uint64_t)((b_base + (uint32_t)(ky * 2 + 0) * B_BLOCK_BYTES) >> 4);
                    const uint64_t blo = b_hi | (uint64_t)((b_base + (uint32_t)(ky * 2 + 1) * B_BLOCK_BYTES) >> 4);
                    const uint32_t acc = (uint32_t)((it | ky) != 0);
                    if (A.x3 == 3) {                                  // plain bf16: first-term product only
#pragma unroll
                        for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(tmem + mb * NB, ahi + (uint64_t)(mb * 128), bhi, idesc_b, acc);
                    } else if (A.x3 == 2) {                           // bf16x3: same planes / blocks, kind::f16, K = 16 per MMA
#pragma unroll
                        for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(tmem + mb * NB, alo + (uint64_t)(mb * 128), bhi, idesc_b, acc);
#pragma unroll
                        for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(tmem + mb * NB, ahi + (uint64_t)(mb * 128), blo, idesc_b, 1u);
#pragma unroll
                        for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(tmem + mb * NB, ahi + (uint64_t)(mb * 128), bhi, idesc_b, 1u);
                    } else {
                        // both cross terms in ONE kind::f16 MMA of K = 16: [lo ; x] (planes 2, 3) x [w ; w - hi] (block ky, 1)
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(tmem + mb * NB, alo + (uint64_t)(mb * 128), blo, idesc_b, acc);
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb) tc::mma_tf32_w(tmem + mb * NB, ahi + (uint64_t)(mb * 128), bhi, idesc, 1u);
                    }
                }
            } else if (A.one) {
                for (int k8 = 0; k8 < nk8; ++k8) {
                    const uint64_t bd = b_hi | (uint64_t)((b_base + (uint32_t)k8 * B_BLOCK_BYTES) >> 4);
                    const uint64_t ad0 = a_hi | (uint64_t)((a_base + (uint32_t)(2 * k8) * PLANE_BYTES + (uint32_t)PITCH * 16u) >> 4);
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb)
                        tc::mma_tf32_w(tmem + mb * NB, ad0 + (uint64_t)(mb * 128), bd, idesc, (uint32_t)((it | k8) != 0));
                }
            } else
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
                for (int k8 = 0; k8 < nk8; ++k8) {
                    if (A.dbg & 1) continue;
                    const uint64_t bd = b_hi | (uint64_t)((b_base + (uint32_t)(ky * 2 + k8) * B_BLOCK_BYTES) >> 4);
                    // address field counts 16-byte units = pixel rows: +128 per M-block
                    const uint64_t ad0 = a_hi | (uint64_t)((a_base + (uint32_t)(2 * k8) * PLANE_BYTES + (uint32_t)(PITCH + (ky - 1) * PITCH) * 16u) >> 4);
                    const uint32_t acc = (uint32_t)((it | ky | k8) != 0);
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb) tc::mma_tf32_w(tmem + mb * NB, ad0 + (uint64_t)(mb * 128), bd, idesc, acc);
                }
            }
            tc::tc_commit_w(bars + 2 + s);
            if (lane == 0) ENDO_TRACE(16 + c * 8 + 7);
        }
        tc::tc_commit_w(bars + 4);
    }
    if (tid == 0) ENDO_TRACE(4);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
    if (tid == 0) { ENDO_TRACE(5); if ((A.dbg & 4) && blockIdx.x == 0 && blockIdx.z == 0) g_tc_trace[0] = nchunks; }
}

}  // namespace tcconv
}  // namespace endo

namespace endo {
// =====================================================================================================
// DenseLayer data gradient on tcgen05 (backward of conv3x3 -> ReLU -> BatchNorm, first term; see the
// lazy-correction scheme in net_kernels.cuh).
//
//   gA[p][ci] = sum_{ky,kx} sum_co G[p + (ky-1, kx-1)][co] * W[co][ci][2-ky][2-kx]
//
// GEMM per 128-pixel M-block: M = 128 linear pixels of the halo tile, N = 64 input channels (one "ci chunk"),
// K = 9 taps x 16 (12 real output channels + zero padding).  The A operand is the output-gradient tile
// (with the lazy BN correction g + A_c + B_c x applied while staging), stored once per CTA in the same
// 16-byte-pitch plane layout as the forward kernel, so BOTH tap offsets are start-address offsets.  The
// accumulator (128 lanes x 64 columns) is double-buffered in TMEM (8 buffers) so the tensor core runs ahead
// of the epilogue.  The epilogue is the memory-bound part (reads x and the gradient buffer, writes the
// gradient buffer: 12 B per pixel-channel): TMEM -> registers -> shared (transposed) so that the global
// read-modify-write is done with lane = channel quad (256 contiguous bytes per pixel) and the per-channel
// BatchNorm-backward sums (sum gy, sum gy*xhat) live in registers of the lane that owns the channel.
// =====================================================================================================
__device__ __forceinline__ uint32_t tcwgrad_pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

namespace tcdgrad {

using tcconv::PITCH; using tcconv::TH; using tcconv::TW; using tcconv::MBLK; using tcconv::PLANE_BYTES;
using tcconv::REAL_ROWS;
constexpr int NC = 64;                                    // input channels per chunk (MMA N)
constexpr int KG = 16;                                    // output-gradient channels incl. padding (MMA K per tap)
constexpr int G_BYTES = (KG / 4) * PLANE_BYTES;           // 78,464
constexpr int WBLK_BYTES = 2 * NC * 16;                   // one (tap, k8) block
constexpr int W_BYTES = 9 * 2 * WBLK_BYTES;               // 36,864
constexpr int TB_PITCH = NC * 4 + 16;                     // 272 B per pixel row (bank spread)
constexpr int TB_BYTES = 128 * TB_PITCH;                  // 34,816
constexpr int NBUF = 8;                                   // TMEM accumulator buffers (8 x 64 = 512 columns)
constexpr int SMEM_BYTES = G_BYTES + W_BYTES + 2 * TB_BYTES + NC * 4 * 4 + 8 * NC * 2 * 4 + 256;
constexpr int NTHREADS = 288;

struct Args {
    const float* g; const float* x; const float* ab;      // gradient buffer, activation buffer, lazy correction [G][C][2]
    const float* coef;                                    // this BN's (a, beta, mean, invstd) [G][Cin][4]
    const float* w;                                       // OIHW [Cout][Cin][3][3]
    const float* wpack;                                   // weight image built by pack_w_dgrad_all_kernel (9216 floats per 64-channel chunk)
    float* gout;                                          // gradient buffer (same as g), accumulated at in_off..in_off+Cin
    float* db;                                            // conv bias gradient [Cout] (sum of the output gradient), accumulated
    double* red; int red_C;                               // [G][red_C][2] BN backward sums
    int C, out_off, Cout, in_off, Cin, H, W, B, G;
    // by-products for the weight-gradient GEMM (net_wgrad3.cuh; nullptr = off), both bf16 and PLANE-MAJOR, [channel group of
    // 8][B*H*W pixels][8 channels] (16 consecutive pixels of a group = 256 contiguous bytes: what a TMA box line wants):
    // relu(bn(x)) of the Cin input channels (the epilogue evaluates it for the ReLU mask) and the corrected output gradient
    // (16 channels, as staged)
    unsigned short* a16; unsigned short* g16;
    // plain mode (TransitionUp, models.py:70-80: the convolution input is the upsampled map, no BatchNorm / ReLU in front):
    // the epilogue stores (first = 1) or accumulates the raw data gradient into a scratch tensor [B,H,W,oC] at channel
    // o_off; up_sum_kernel then folds the 2x2 blocks into the half-resolution gradient buffer
    int plain, first, oC, o_off;
    float* po;
};

__global__ void __launch_bounds__(NTHREADS, 1)
dense_dgrad_tf32_kernel(const Args A) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* g_s = smem;
    unsigned char* w_s = smem + G_BYTES;
    unsigned char* tb0 = smem + G_BYTES + W_BYTES;                          // transposed epilogue tile, DOUBLE-BUFFERED: one barrier per
                                                                           // unit instead of two (r2 ncu: 18 % of the stall samples were barriers)
    float* ctab = reinterpret_cast<float*>(tb0 + 2 * TB_BYTES);            // [NC][4] a, beta, mean, invstd
    float* red = ctab + NC * 4;                                            // [8][NC][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8 * NC * 2);        // w_full, acc_full[8], acc_empty[8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * NBUF);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (A.W + TW - 1) / TW;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    const int b = blockIdx.z;
    const int g = b / (A.B / A.G);
    const int nchunks = (A.Cin + NC - 1) / NC;
    // gridDim.y > 1 (low-resolution levels: fewer tiles than SMs): one 64-channel chunk per CTA -- the chunks write disjoint
    // gradient channels, only the small output-gradient tile is staged once per chunk instead of once per tile
    const int c_begin = gridDim.y > 1 ? (int)blockIdx.y : 0;
    const int c_end = gridDim.y > 1 ? c_begin + 1 : nchunks;

    pdl_trigger();
    if (warp == 8) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        tc::mbar_init(bars + 0, 256);
        for (int i = 0; i < NBUF; ++i) { tc::mbar_init(bars + 1 + i, 1); tc::mbar_init(bars + 1 + NBUF + i, 128); }
        tc::fence_mbar_init();
    }
    pdl_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 8) {
        // ---------------------------------------------------------------- stage the output-gradient tile once
        {
            const int grp = tid & 3;
            const int ch = grp * 4;
            const bool ch_ok = ch < A.Cout;
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
            if (ch_ok) {
                const float* abp = A.ab + ((size_t)g * A.C + A.out_off + ch) * 2;
                c0 = __ldg(reinterpret_cast<const float4*>(abp));
                c1 = __ldg(reinterpret_cast<const float4*>(abp + 4));
            }
            const size_t img = (size_t)b * A.H * A.W;
            float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int part = 0; part < 2; ++part) {                    // 19 pixels per thread in two batches of 10 (2 loads each)
                float4 gq[10], xq[10];
                unsigned okmask = 0u;
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const int px = (tid >> 2) + 64 * (part * 10 + j);
                    const int r = px / PITCH, cc = px - r * PITCH;
                    const int y = y0 + r - 1, x = x0 + cc - 1;
                    gq[j] = make_float4(0.f, 0.f, 0.f, 0.f); xq[j] = gq[j];
                    if (ch_ok && px < REAL_ROWS && y >= 0 && y < A.H && x >= 0 && x < A.W) {
                        const size_t o = (img + (size_t)y * A.W + x) * A.C + A.out_off + ch;
                        gq[j] = __ldg(reinterpret_cast<const float4*>(A.g + o));
                        xq[j] = __ldg(reinterpret_cast<const float4*>(A.x + o));
                        okmask |= 1u << j;
                    }
                }
#pragma unroll
                for (int j = 0; j < 10; ++j) {
                    const int px = (tid >> 2) + 64 * (part * 10 + j);
                    if (px < REAL_ROWS) {
                        const int r = px / PITCH, cc = px - r * PITCH;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (okmask & (1u << j)) {
                            v.x = gq[j].x + fmaf(c0.y, xq[j].x, c0.x); v.y = gq[j].y + fmaf(c0.w, xq[j].y, c0.z);
                            v.z = gq[j].z + fmaf(c1.y, xq[j].z, c1.x); v.w = gq[j].w + fmaf(c1.w, xq[j].w, c1.z);
                        }
                        if (r >= 1 && r <= TH && cc >= 1 && cc <= TW) {          // interior: each pixel belongs to exactly one tile
                            bs.x += v.x; bs.y += v.y; bs.z += v.z; bs.w += v.w;
                            const int y = y0 + r - 1, x = x0 + cc - 1;
                            if (A.g16 && blockIdx.y == 0 && y < A.H && x < A.W)   // (padding channels 12 .. 15: zeros)
                                *reinterpret_cast<uint2*>(A.g16 + (((size_t)(ch >> 3) * A.B * A.H * A.W) + img + (size_t)y * A.W + x) * 8 + (ch & 7)) =
                                    make_uint2(tcwgrad_pack_bf16(v.x, v.y), tcwgrad_pack_bf16(v.z, v.w));
                        }
                        if (okmask & (1u << j)) {
                            v.x = tcconv::tf32_rn(v.x); v.y = tcconv::tf32_rn(v.y); v.z = tcconv::tf32_rn(v.z); v.w = tcconv::tf32_rn(v.w);
                        }
                        *reinterpret_cast<float4*>(g_s + grp * PLANE_BYTES + (size_t)(px + 1) * 16) = v;   // row 0 is a margin row
                    }
                }
            }
            // conv bias gradient = sum of the output gradient over the tile interior (each pixel is interior to one tile)
#pragma unroll
            for (int o2 = 4; o2 < 32; o2 <<= 1) {
                bs.x += __shfl_xor_sync(0xffffffffu, bs.x, o2); bs.y += __shfl_xor_sync(0xffffffffu, bs.y, o2);
                bs.z += __shfl_xor_sync(0xffffffffu, bs.z, o2); bs.w += __shfl_xor_sync(0xffffffffu, bs.w, o2);
            }
            if (A.db && lane < 4 && ch_ok && blockIdx.y == 0) {
                atomicAdd(A.db + ch, bs.x); atomicAdd(A.db + ch + 1, bs.y);
                atomicAdd(A.db + ch + 2, bs.z); atomicAdd(A.db + ch + 3, bs.w);
            }
        }
        const int q = warp & 3, chalf = warp >> 2;            // TMEM lane quadrant, column half (32 columns)
        const int quad = lane & 15, psub = lane >> 4;         // phase-2 mapping: channel quad, pixel parity
        int unit = 0;
        for (int c = c_begin; c < c_end; ++c) {
            const int ci0 = c * NC;
            // ------------------------------------------------------------ weights + BN table of this ci chunk
            {   // 36,864-byte weight image of this chunk, copied verbatim (9 x 16 B per thread, all loads first)
                const float4* src = reinterpret_cast<const float4*>(A.wpack + (size_t)c * 9216);
                float4 wq[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) wq[j] = __ldg(src + tid + 256 * j);
#pragma unroll
                for (int j = 0; j < 9; ++j) *reinterpret_cast<float4*>(w_s + (size_t)(tid + 256 * j) * 16) = wq[j];
            }
            if (tid < NC) {
                const int ci = ci0 + tid;
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ci < A.Cin && !A.plain) e = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)g * A.Cin + ci) * 4));
                *reinterpret_cast<float4*>(ctab + tid * 4) = e;
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bars + 0);                         // w_full: phase c
            asm volatile("bar.sync 1, 256;" ::: "memory");     // ctab visible to all epilogue threads
            // per-lane constants for the 4 channels this lane owns in phase 2
            float ca[4], cb[4], cm[4], cs[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 t4 = *reinterpret_cast<const float4*>(ctab + (quad * 4 + e) * 4);
                ca[e] = t4.x; cb[e] = t4.y; cm[e] = t4.z; cs[e] = t4.w;
            }
            const bool quad_ok = (ci0 + quad * 4) < A.Cin;
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            for (int mb = 0; mb < MBLK; ++mb, ++unit) {
                const int buf = unit % NBUF;
                unsigned char* tb = tb0 + (unit & 1) * TB_BYTES;
                // the activations / gradient rows this lane will update (phase 2) do not depend on the accumulator: request them
                // first, so that they travel while the MMAs finish, the accumulator is drained and the warps meet at the barrier
                float4 xv[8];
                unsigned okmask = 0u;
                unsigned pix[8];                                          // linear pixel index (b * H + y) * W + x
                const size_t npix_all = (size_t)A.B * A.H * A.W;
                const size_t cbase = (A.plain ? A.o_off : A.in_off) + ci0 + quad * 4;
                const size_t cstr = A.plain ? A.oC : A.C;
                if (quad_ok) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {                     // all 8 loads of this unit in flight at once
                        const int p = warp * 16 + it * 2 + psub;
                        const int L = PITCH + mb * 128 + p;
                        const int r = L / PITCH, cc = L - r * PITCH;
                        const int y = y0 + r - 1, x = x0 + cc - 1;
                        pix[it] = 0;
                        if ((r <= TH) && (cc >= 1) && (cc <= TW) && (y < A.H) && (x < A.W)) {
                            pix[it] = (unsigned)((b * A.H + y) * A.W + x);
                            if (!A.plain) xv[it] = __ldg(reinterpret_cast<const float4*>(A.x + (size_t)pix[it] * cstr + cbase));
                            okmask |= 1u << it;
                        }
                    }
                }
                tc::mbar_wait(bars + 1 + buf, (unit / NBUF) & 1);
                tc::tc_fence_after();
                // phase 1: TMEM -> registers -> transposed shared tile [pixel][channel]
                {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * NC + chalf * 32;
                    float v[16];
                    unsigned char* row = tb + (size_t)(q * 32 + lane) * TB_PITCH + chalf * 128;
                    tc::tmem_ld16(taddr, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(row + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    tc::tmem_ld16(taddr + 16, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(row + 64 + j * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                tc::tc_fence_before();
                asm volatile("bar.sync 1, 256;" ::: "memory");               // all 8 warps have drained their TMEM part
                if (chalf == 0) tc::mbar_arrive(bars + 1 + NBUF + buf);      // 128 arrivals: accumulator buffer free again
                // phase 2: lane = (pixel parity, channel quad); 16 pixels per warp
                if (quad_ok) {
                    if (A.plain) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            if (okmask & (1u << it)) {
                                const int p = warp * 16 + it * 2 + psub;
                                const float4 d = *reinterpret_cast<const float4*>(tb + (size_t)p * TB_PITCH + quad * 16);
                                if (A.first) *reinterpret_cast<float4*>(A.po + (size_t)pix[it] * cstr + cbase) = d;
                                else tcconv::red_add_v4(A.po + (size_t)pix[it] * cstr + cbase, d.x, d.y, d.z, d.w);
                            }
                        }
                    } else
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        if (okmask & (1u << it)) {
                            const int p = warp * 16 + it * 2 + psub;
                            const float4 d = *reinterpret_cast<const float4*>(tb + (size_t)p * TB_PITCH + quad * 16);
                            const float4 xq = xv[it];
                            const float e0 = xq.x - cm[0], e1 = xq.y - cm[1], e2 = xq.z - cm[2], e3 = xq.w - cm[3];
                            const float t0 = fmaf(ca[0], e0, cb[0]), t1 = fmaf(ca[1], e1, cb[1]);
                            const float t2 = fmaf(ca[2], e2, cb[2]), t3 = fmaf(ca[3], e3, cb[3]);
                            const float g0 = t0 > 0.f ? d.x : 0.f;
                            const float g1 = t1 > 0.f ? d.y : 0.f;
                            const float g2 = t2 > 0.f ? d.z : 0.f;
                            const float g3 = t3 > 0.f ? d.w : 0.f;
                            s1[0] += g0; s2[0] += g0 * (e0 * cs[0]);
                            s1[1] += g1; s2[1] += g1 * (e1 * cs[1]);
                            s1[2] += g2; s2[2] += g2 * (e2 * cs[2]);
                            s1[3] += g3; s2[3] += g3 * (e3 * cs[3]);
                            // gout[p][ci] += a * g: each (pixel, channel) is touched by exactly one thread of one CTA per launch
                            tcconv::red_add_v4(A.gout + (size_t)pix[it] * cstr + cbase, ca[0] * g0, ca[1] * g1, ca[2] * g2, ca[3] * g3);
                            if (A.a16) {                                 // relu(bn(x)) for the weight-gradient GEMM
                                unsigned short* ap = A.a16 + ((size_t)((ci0 >> 3) + (quad >> 1)) * npix_all + pix[it]) * 8 + (quad & 1) * 4;
                                *reinterpret_cast<uint2*>(ap) = make_uint2(tcwgrad_pack_bf16(fmaxf(t0, 0.f), fmaxf(t1, 0.f)),
                                                                           tcwgrad_pack_bf16(fmaxf(t2, 0.f), fmaxf(t3, 0.f)));
                                if (ci0 + quad * 4 + 4 == A.Cin && (A.Cin & 4)) *reinterpret_cast<uint2*>(ap + 4) = make_uint2(0u, 0u);
                            }
                        }
                    }
                }
                // (no barrier here: the next unit writes the OTHER transposed tile; this one is rewritten two units later, after the
                // barrier of the next unit, which every thread reaches only when it has finished reading this one)
            }
            // ------------------------------------------------------------ BN-backward sums of this chunk
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                if (psub == 0) {
                    red[(warp * NC + quad * 4 + e) * 2] = s1[e];
                    red[(warp * NC + quad * 4 + e) * 2 + 1] = s2[e];
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < 2 * NC) {
                const int j = tid >> 1, which = tid & 1;
                if (ci0 + j < A.Cin && !A.plain) {
                    double sum = 0.0;
#pragma unroll
                    for (int wq = 0; wq < 8; ++wq) sum += (double)red[(wq * NC + j) * 2 + which];
                    atomicAdd(A.red + ((size_t)g * A.red_C + ci0 + j) * 2 + which, sum);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");               // red / ctab / weights reusable
        }
    } else {
        // -------------------------------------------------------------------- MMA issuer: warp 8, convergent; one elected lane issues
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t idesc = tc::instr_desc(tc::FMT_TF32, 128, NC);
        const uint32_t g_base = tc::smem_u32(g_s), w_base = tc::smem_u32(w_s);
        int unit = 0;
        for (int c = c_begin; c < c_end; ++c) {
            tc::mbar_wait(bars + 0, (c - c_begin) & 1);
            tc::tc_fence_after();
            // groups of up to 4 M-blocks are issued interleaved (independent accumulators: a dependent MMA waits
            // ~266 cycles for its predecessor); the other 4 TMEM buffers belong to the group the epilogue is draining
            for (int mb0 = 0; mb0 < MBLK; mb0 += 4) {
                const int gsz = (MBLK - mb0) < 4 ? (MBLK - mb0) : 4;
                for (int u = 0; u < gsz; ++u) {
                    const int un = unit + u, buf = un % NBUF;
                    if (un >= NBUF) tc::mbar_wait(bars + 1 + NBUF + buf, ((un / NBUF) - 1) & 1);
                }
                tc::tc_fence_after();
                const uint64_t a_hi = tc::smem_desc(0, PLANE_BYTES, 128), b_hi = tc::smem_desc(0, NC * 16, 128);
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap - 3 * ky;
#pragma unroll
                    for (int k8 = 0; k8 < 2; ++k8) {
                        const uint64_t bd = b_hi | (uint64_t)((w_base + (uint32_t)(tap * 2 + k8) * WBLK_BYTES) >> 4);
                        const uint64_t ad0 = a_hi | (uint64_t)((g_base + (uint32_t)(2 * k8) * PLANE_BYTES +
                                                               (uint32_t)(1 + PITCH + (ky - 1) * PITCH + (kx - 1)) * 16u) >> 4);
                        for (int u = 0; u < gsz; ++u) {
                            const int buf = (unit + u) % NBUF;
                            tc::mma_tf32_w(tmem + buf * NC, ad0 + (uint64_t)((mb0 + u) * 128), bd, idesc, (uint32_t)((tap | k8) != 0));
                        }
                    }
                }
                for (int u = 0; u < gsz; ++u) tc::tc_commit_w(bars + 1 + (unit + u) % NBUF);
                unit += gsz;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
}

// low[p2][c] += sum over the 2x2 block of hi[p][c]  (backward of nearest-neighbour upsampling, models.py:73); thread =
// (coarse pixel, channel quad)
__global__ void __launch_bounds__(256)
up_sum_kernel(const float* __restrict__ hi, int hiC, float* __restrict__ low, int lowC, int low_off, int B, int h2, int w2, int C) {
    pdl_enter();
    const int nq = C >> 2;
    const long long total = (long long)B * h2 * w2 * nq;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int q = (int)(i % nq);
        const long long pp = i / nq;
        const int x2 = (int)(pp % w2), y2 = (int)((pp / w2) % h2), b = (int)(pp / ((long long)w2 * h2));
        const float* p00 = hi + (((size_t)b * 2 * h2 + 2 * y2) * 2 * w2 + 2 * x2) * hiC + q * 4;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p00)), b4 = __ldg(reinterpret_cast<const float4*>(p00 + hiC));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(p00 + (size_t)2 * w2 * hiC));
        const float4 d4 = __ldg(reinterpret_cast<const float4*>(p00 + (size_t)2 * w2 * hiC + hiC));
        float4* o = reinterpret_cast<float4*>(low + (size_t)pp * lowC + low_off + q * 4);
        float4 v = *o;
        v.x += (a.x + b4.x) + (c4.x + d4.x); v.y += (a.y + b4.y) + (c4.y + d4.y);
        v.z += (a.z + b4.z) + (c4.z + d4.z); v.w += (a.w + b4.w) + (c4.w + d4.w);
        *o = v;
    }
}

}  // namespace tcdgrad

// =====================================================================================================
// DenseLayer weight gradient on tcgen05 (bf16 operands, fp32 accumulation in TMEM).
//
//   dW[co][ci][ky][kx] = sum_p G[p][co] * act[p + (ky-1, kx-1)][ci],   act = relu(bn(x)), zero outside the image
//
// The reduction runs over PIXELS, so pixels are the GEMM K dimension and both operands are "MN-major": for a
// group of 8 channels, 8 consecutive pixels x 16 bytes form one 128-byte core matrix -- which is exactly the
// plane layout of the forward kernel (pixels 16 B apart), with bf16 packing 8 channels into the 16 bytes.
//   A = act planes          M = 64 input channels of this CTA's channel block (rows 64..127 of the MMA read
//                           whatever follows in shared memory and are ignored), K = 16 pixels.  ONLY THE 8x32 TILE
//                           INTERIOR is staged (the two pad columns of the pitch-34 rows stay zero): the halo a 3x3
//                           tap needs sits on the small operand,
//   B = shifted G planes    N = 3 (kx) x 16 (co): the (8+2) x 34 halo tile of the OUTPUT GRADIENT (real values of the
//                           neighbouring tiles, zero outside the image; 12 channels instead of Cin) with the horizontal
//                           taps folded into N by writing each staged gradient pixel into three planes shifted by kx
//                           rows; the vertical tap is a B start-address offset of -+34 rows:
//                               dW[ky][kx] = sum_{p in tile} act[p] * G[p - (ky-1, kx-1)]
//                           (round 1 shifted the activations instead and read 10/8 of them plus 64-channel padding:
//                           1.59x the algorithmic DRAM bytes, ncu).  Three accumulators D_ky[ci][kx*16+co] (144 TMEM
//                           columns) stay resident while the CTA walks over its share of 8x32-pixel tiles; one
//                           fp32 atomicAdd per weight at the end.
// =====================================================================================================
namespace tcwgrad {

constexpr int PITCH = 34, TR = 8, TW = 32;
constexpr int KPX = TR * PITCH;                          // 272 pixels = 17 K-steps of 16
constexpr int A_ROWS = (TR + 2) * PITCH;                 // 340 staged activation pixels
constexpr int PLANE_BYTES = 345 * 16;                    // 5,520: = 16 (mod 128) so the 8 channel groups of one pixel spread over all banks
constexpr int MCH = 64;                                  // input channels per CTA block
constexpr int A_STAGE = (MCH / 8) * PLANE_BYTES;         // 44,160
constexpr int G_STAGE = 6 * PLANE_BYTES;                 // 33,120   planes [kx][co half], one margin row in front
constexpr int STAGE = A_STAGE + G_STAGE;                 // 77,280
constexpr int KTAB_BYTES = 2 * 8 * 144;                  // BatchNorm coefficients of this CTA's 64 channels, both statistic groups:
                                                         // [g][8-channel group][8 x float4 + 16 B pad] (144 B rows: the 8 groups a warp reads hit distinct banks)
constexpr int SMEM_BYTES = 2 * STAGE + 8 * PLANE_BYTES + KTAB_BYTES + 256;   // tail pad: the M = 128 read of stage 1 stays inside the allocation
constexpr int NB = 48;
constexpr int NPROD = 512;                               // 16 producer warps: the staging is ALU work (BN + ReLU + bf16 packing, ~3000
                                                         // instructions per thread and tile with 8 warps = 2 warps per scheduler: dependent-issue bound, ncu)
constexpr int NTHREADS = NPROD + 32;

struct Args {
    const float* x; const float* coef;                   // activations + this BN's (a, beta, mean, invstd) [G][Cin][4]
    const float* g; const float* ab;                     // gradient buffer + lazy correction [G][C][2]
    float* dw;                                           // OIHW gradient (accumulated); the bias gradient comes from the dgrad kernel
    int C, in_off, Cin, out_off, Cout, H, W, B, G;
    int tiles_per_cta, n_tiles;
    const float* xa; int xa_C, up;                       // activation source: buffer + channel stride; up = 1: half-resolution buffer,
                                                         // nearest-upsampled x2, no BatchNorm (TransitionUp)
    // 1x1 mode (TransitionDown, models.py:56-67): no taps, N = 48 output channels [out_off, out_off + 48) per launch; the
    // output gradient is the max-pool-routed gradient of the NEXT level's buffer (gc, xc, abc: stride cC, first channel c_off)
    int one; const unsigned char* argmax; const float* gc; const float* xc; const float* abc; int cC, c_off, cH, cW;
    int dbg;                                             // ENDO_TC_DEBUG bit 8: clock64 trace of CTA (0, 0) (tools/trace_wgrad.py)
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__global__ void __launch_bounds__(NTHREADS, 1)
dense_wgrad_bf16_kernel(const Args A) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* ktab = reinterpret_cast<float*>(smem + 2 * STAGE + 8 * PLANE_BYTES);         // [G <= 2][MCH][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE + 8 * PLANE_BYTES + KTAB_BYTES);   // full[2], empty[2], accum
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.y * MCH;
    const int tiles_x = (A.W + TW - 1) / TW, tiles_y = (A.H + TR - 1) / TR;
    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    const int ntiles = t_end - t_begin;

    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    if (tid < 2 * MCH) {                                     // coefficient table (zeros for TransitionUp: no BatchNorm in front)
        const int gi = tid / MCH, cl = tid % MCH, ch = ci0 + cl;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < A.G && ch < A.Cin && !A.up) e = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)gi * A.Cin + ch) * 4));
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(ktab) + (gi * 8 + (cl >> 3)) * 144 + (cl & 7) * 16) = e;
    }
    if (tid == 0) {
        tc::mbar_init(bars + 0, NPROD); tc::mbar_init(bars + 1, NPROD);
        tc::mbar_init(bars + 2, 1);   tc::mbar_init(bars + 3, 1);
        tc::mbar_init(bars + 4, 1);
        tc::fence_mbar_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 16) {
        // Both stages are cleared ONCE: every tile writes the same rows (activation planes: the 8x32 interior, a pixel
        // outside the image stores zeros; gradient planes: rows kx .. 339 + kx of plane kx), every other row -- the pad columns
        // of the activation rows, the margins of the kx-shifted planes -- stays zero.
        for (int i = tid; i < 2 * (STAGE / 16); i += NPROD) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1;
            if (it >= 2) tc::mbar_wait(bars + 2 + s, ((it >> 1) - 1) & 1);
            const int t = t_begin + it;
            const int b = t / (tiles_x * tiles_y), rem = t - b * (tiles_x * tiles_y);
            const int tx = rem / tiles_y, ty = rem - tx * tiles_y;      // column-major: consecutive tiles of a CTA are vertical
                                                                        // neighbours, the 2 shared halo rows hit L2 instead of DRAM
            const int y0 = ty * TR, x0 = tx * TW;
            const int g = b / (A.B / A.G);
            unsigned char* a_s = smem + s * STAGE;
            unsigned char* g_s = a_s + A_STAGE;
            const size_t img = (size_t)b * A.H * A.W;
            // ---- output gradient, round 0: the loads are issued FIRST and consumed after the activation planes are written
            //      (r2 ncu: three serialised load round trips per tile were 37 % of all warp-stall samples of this kernel)
            float4 gq0[2], xq0[2];
            bool g0_ok = false;
            int g0_q = 0, g0_hf = 0;
            size_t g0_oo = 0;
            if (!A.one) {
                g0_hf = tid >= A_ROWS ? 1 : 0; g0_q = tid - g0_hf * A_ROWS;
                const int r = g0_q / PITCH, cc = g0_q - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                g0_ok = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                g0_oo = (img + (size_t)y * A.W + x) * A.C + A.out_off + g0_hf * 8;
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    gq0[h4] = make_float4(0.f, 0.f, 0.f, 0.f); xq0[h4] = gq0[h4];
                    if (g0_ok && g0_hf * 8 + h4 * 4 < A.Cout) {
                        gq0[h4] = __ldg(reinterpret_cast<const float4*>(A.g + g0_oo + h4 * 4));
                        xq0[h4] = __ldg(reinterpret_cast<const float4*>(A.x + g0_oo + h4 * 4));
                    }
                }
            }
            // ---- activations: (pixel, 8-channel group) items, BN+ReLU, bf16
            {
                const int grp = tid & 7;
                const int ch = ci0 + grp * 8;
                const bool ch_ok = ch < A.Cin;                       // Cin is a multiple of 4: a group may be half valid
                const bool hi_ok = ch + 4 < A.Cin;
                // (a, beta, mean, invstd) of this thread's 8 channels: registers for the whole tile
                const float4* kt = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(ktab) + (g * 8 + grp) * 144);
                const float4 k0 = kt[0], k1 = kt[1], k2 = kt[2], k3 = kt[3], k4 = kt[4], k5 = kt[5], k6 = kt[6], k7 = kt[7];
                const int sh = A.up ? 1 : 0;
                const int sW = A.W >> sh;
                const float* xa_b = A.xa + (size_t)b * (A.H >> sh) * sW * A.xa_C + A.in_off + ch;
                const bool v8_ok = (((A.in_off + ci0) | A.xa_C) & 7) == 0 && (reinterpret_cast<uintptr_t>(A.xa) & 31) == 0;
                float4 q0[4], q1[4];                                 // 4 interior pixels per thread, all loads first
                unsigned okmask = 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ip = (tid >> 3) + 64 * j;              // interior pixel 0 .. 255
                    const int y = y0 + (ip >> 5), x = x0 + (ip & 31);
                    q0[j] = make_float4(0.f, 0.f, 0.f, 0.f); q1[j] = q0[j];
                    if (ch_ok && y < A.H && x < A.W) {
                        const float* p = xa_b + ((size_t)(y >> sh) * sW + (x >> sh)) * A.xa_C;
                        if (hi_ok && v8_ok) {
                            // one 256-bit load = one whole 32-byte sector per lane, not cached in the (small) L1
                            asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                         : "=f"(q0[j].x), "=f"(q0[j].y), "=f"(q0[j].z), "=f"(q0[j].w),
                                           "=f"(q1[j].x), "=f"(q1[j].y), "=f"(q1[j].z), "=f"(q1[j].w) : "l"(p));
                        } else {
                            q0[j] = __ldg(reinterpret_cast<const float4*>(p));
                            if (hi_ok) q1[j] = __ldg(reinterpret_cast<const float4*>(p + 4));
                        }
                        okmask |= 1u << j;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ip = (tid >> 3) + 64 * j;
                    const int px = (1 + (ip >> 5)) * PITCH + 1 + (ip & 31);      // row of the pitch-34 plane
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (okmask & (1u << j)) {
                        const float4 a0 = q0[j], a1 = q1[j];
                        float v0 = a0.x, v1 = a0.y, v2 = a0.z, v3 = a0.w, v4 = a1.x, v5 = a1.y, v6 = a1.z, v7 = a1.w;
                        if (!A.up) {
                            v0 = fmaxf(fmaf(k0.x, a0.x - k0.z, k0.y), 0.f); v1 = fmaxf(fmaf(k1.x, a0.y - k1.z, k1.y), 0.f);
                            v2 = fmaxf(fmaf(k2.x, a0.z - k2.z, k2.y), 0.f); v3 = fmaxf(fmaf(k3.x, a0.w - k3.z, k3.y), 0.f);
                            v4 = v5 = v6 = v7 = 0.f;
                            if (hi_ok) {
                                v4 = fmaxf(fmaf(k4.x, a1.x - k4.z, k4.y), 0.f); v5 = fmaxf(fmaf(k5.x, a1.y - k5.z, k5.y), 0.f);
                                v6 = fmaxf(fmaf(k6.x, a1.z - k6.z, k6.y), 0.f); v7 = fmaxf(fmaf(k7.x, a1.w - k7.z, k7.y), 0.f);
                            }
                        }
                        o = make_uint4(pack_bf16(v0, v1), pack_bf16(v2, v3), pack_bf16(v4, v5), pack_bf16(v6, v7));
                    }
                    *reinterpret_cast<uint4*>(a_s + grp * PLANE_BYTES + (size_t)px * 16) = o;
                }
            }
            // ---- output gradient: plane (kx, half) row (1 + q) holds G[q - (kx-1)][half*8 .. +8], zero outside the tile
            //      interior.  Two threads per interior pixel (one per 8-channel half): all loads are issued before any use,
            //      then the corrected channels are written into the three kx-shifted planes.
            const int gpix = tid & 255, half = tid >> 8;
            if (A.one) {
                // 1x1 mode: 48 channels in three rounds of 16 (argmax word + g + x loads first)
                const int r = 1 + (gpix >> 5), cc = 1 + (gpix & 31);
                const int y = y0 + r - 1, x = x0 + cc - 1;
                const bool ok = (y < A.H) && (x < A.W);
                const unsigned pos = (unsigned)(((y & 1) << 1) | (x & 1));
                const size_t pp = ok ? ((size_t)(b * A.cH + (y >> 1)) * A.cW + (x >> 1)) : 0;
                const int q = r * PITCH + cc;
#pragma unroll 1
                for (int sub = 0; sub < 3; ++sub) {
                    const int cbase = A.out_off + sub * 16 + half * 8;      // channel inside the conv's output
                    unsigned am[2];
                    float4 gq[2], xq[2];
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        am[h4] = 0xffffffffu; gq[h4] = make_float4(0.f, 0.f, 0.f, 0.f); xq[h4] = gq[h4];
                        if (ok && cbase + h4 * 4 < A.Cout) {
                            am[h4] = __ldg(reinterpret_cast<const unsigned*>(A.argmax + pp * A.Cout + cbase + h4 * 4));
                            gq[h4] = __ldg(reinterpret_cast<const float4*>(A.gc + pp * A.cC + A.c_off + cbase + h4 * 4));
                            xq[h4] = __ldg(reinterpret_cast<const float4*>(A.xc + pp * A.cC + A.c_off + cbase + h4 * 4));
                        }
                    }
                    float v[8];
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                        if (cbase + h4 * 4 < A.Cout) {
                            const float* abp = A.abc + ((size_t)g * A.cC + A.c_off + cbase + h4 * 4) * 2;
                            c0 = __ldg(reinterpret_cast<const float4*>(abp)); c1 = __ldg(reinterpret_cast<const float4*>(abp + 4));
                        }
                        v[h4 * 4 + 0] = ((am[h4] & 0xffu) == pos) ? gq[h4].x + fmaf(c0.y, xq[h4].x, c0.x) : 0.f;
                        v[h4 * 4 + 1] = (((am[h4] >> 8) & 0xffu) == pos) ? gq[h4].y + fmaf(c0.w, xq[h4].y, c0.z) : 0.f;
                        v[h4 * 4 + 2] = (((am[h4] >> 16) & 0xffu) == pos) ? gq[h4].z + fmaf(c1.y, xq[h4].z, c1.x) : 0.f;
                        v[h4 * 4 + 3] = ((am[h4] >> 24) == pos) ? gq[h4].w + fmaf(c1.w, xq[h4].w, c1.z) : 0.f;
                    }
                    const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    *reinterpret_cast<uint4*>(g_s + (sub * 2 + half) * PLANE_BYTES + (size_t)(q + 1) * 16) = ok ? o : make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                // (8 + 2) x 34 halo tile, two threads per pixel (one per 8-channel half): 680 items in two rounds
#pragma unroll 1
                for (int rnd = 0; rnd < 2; ++rnd) {
                    const int i = tid + NPROD * rnd;
                    if (i >= 2 * A_ROWS) break;
                    int hf = g0_hf, q = g0_q;
                    bool ok = g0_ok;
                    float4 gq[2] = {gq0[0], gq0[1]}, xq[2] = {xq0[0], xq0[1]};
                    if (rnd == 1) {
                        hf = i >= A_ROWS ? 1 : 0; q = i - hf * A_ROWS;
                        const int r = q / PITCH, cc = q - r * PITCH;
                        const int y = y0 + r - 1, x = x0 + cc - 1;
                        ok = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                        const size_t oo = (img + (size_t)y * A.W + x) * A.C + A.out_off + hf * 8;
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            gq[h4] = make_float4(0.f, 0.f, 0.f, 0.f); xq[h4] = gq[h4];
                            if (ok && hf * 8 + h4 * 4 < A.Cout) {
                                gq[h4] = __ldg(reinterpret_cast<const float4*>(A.g + oo + h4 * 4));
                                xq[h4] = __ldg(reinterpret_cast<const float4*>(A.x + oo + h4 * 4));
                            }
                        }
                    }
                    float v[8];
                    const float* abp = A.ab + ((size_t)g * A.C + A.out_off + hf * 8) * 2;
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        if (ok && hf * 8 + h4 * 4 < A.Cout) {
                            const float4 c0 = __ldg(reinterpret_cast<const float4*>(abp + h4 * 8));
                            const float4 c1 = __ldg(reinterpret_cast<const float4*>(abp + h4 * 8 + 4));
                            v[h4 * 4 + 0] = gq[h4].x + fmaf(c0.y, xq[h4].x, c0.x); v[h4 * 4 + 1] = gq[h4].y + fmaf(c0.w, xq[h4].y, c0.z);
                            v[h4 * 4 + 2] = gq[h4].z + fmaf(c1.y, xq[h4].z, c1.x); v[h4 * 4 + 3] = gq[h4].w + fmaf(c1.w, xq[h4].w, c1.z);
                        } else {
                            v[h4 * 4 + 0] = v[h4 * 4 + 1] = v[h4 * 4 + 2] = v[h4 * 4 + 3] = 0.f;
                        }
                    }
                    const uint4 o = ok ? make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]))
                                       : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        *reinterpret_cast<uint4*>(g_s + (kx * 2 + hf) * PLANE_BYTES + (size_t)(q + kx) * 16) = o;
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bars + s);
        }
        // ---- epilogue: D_ky[ci][kx*16 + co] -> atomicAdd into OIHW
        tc::mbar_wait(bars + 4, 0);
        tc::tc_fence_after();
        if (warp < 2 && ntiles > 0 && A.one) {
            const int ci = ci0 + warp * 32 + lane;
#pragma unroll 1
            for (int grp16 = 0; grp16 < 3; ++grp16) {
                float acc16[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc16[j] = 0.f;
#pragma unroll 1
                for (int set = 0; set < 9; ++set) {
                    float v[16];
                    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + set * NB + grp16 * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc16[j] += v[j];
                }
                if (ci < A.Cin) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = A.out_off + grp16 * 16 + j;
                        if (co < A.Cout) atomicAdd(A.dw + (size_t)co * A.Cin + ci, acc16[j]);
                    }
                }
            }
        } else if (warp < 2 && ntiles > 0) {
            const int ci = ci0 + warp * 32 + lane;
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
                for (int kx = 0; kx < 3; ++kx) {
                    float v[16], v1[16], v2[16];                      // the three interleaved accumulator sets
                    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + ky * NB + kx * 16, v);
                    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (3 + ky) * NB + kx * 16, v1);
                    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (6 + ky) * NB + kx * 16, v2);
#pragma unroll
                    for (int co = 0; co < 16; ++co) v[co] += v1[co] + v2[co];
                    if (ci < A.Cin) {
#pragma unroll
                        for (int co = 0; co < 16; ++co)
                            if (co < A.Cout) atomicAdd(A.dw + (((size_t)co * A.Cin + ci) * 3 + ky) * 3 + kx, v[co]);
                    }
                }
            }
        }
    } else {                                                 // warp 16, convergent; one elected lane issues
        const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t idesc = tc::instr_desc(tc::FMT_BF16, 128, NB, 1, 1);
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1;
            tc::mbar_wait(bars + s, (it >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t a_base = tc::smem_u32(smem + s * STAGE);
            const uint32_t g_base = a_base + A_STAGE;
            // consecutive K-steps rotate over three accumulator sets (9 independent chains): an MMA that accumulates
            // into the tile its predecessor wrote waits ~266 cycles for it
            const uint64_t d_hi = tc::smem_desc(0, 128, PLANE_BYTES);   // MN-major: LBO = 8-pixel groups (128 B), SBO = 8-channel groups (planes)
            const uint64_t a_d0 = d_hi | (uint64_t)(a_base >> 4), b_d0 = d_hi | (uint64_t)((g_base + 16u) >> 4);
            if (A.one) {
                // 1x1: one MMA per 16 pixels, rotating over 9 accumulator tiles
#pragma unroll 1
                for (int k16 = 0; k16 < KPX / 16; ++k16)
                    tc::mma_f16_w(tmem + (k16 % 9) * NB, a_d0 + (uint64_t)(PITCH + k16 * 16), b_d0 + (uint64_t)(PITCH + k16 * 16), idesc,
                                (uint32_t)(it != 0 || k16 >= 9));
            } else {
#pragma unroll 1
            for (int k16 = 0; k16 < KPX / 16; ++k16) {
                const int set = k16 % 3;
                const uint64_t ad = a_d0 + (uint64_t)(PITCH + k16 * 16);
                const uint32_t acc = (uint32_t)(it != 0 || k16 >= 3);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)          // act[p] pairs with G[p - (ky-1) rows]: B rows slide, A stays
                    tc::mma_f16_w(tmem + (set * 3 + ky) * NB, ad, b_d0 + (uint64_t)(PITCH + k16 * 16 - (ky - 1) * PITCH), idesc, acc);
            }
            }
            tc::tc_commit_w(bars + 2 + s);
        }
        tc::tc_commit_w(bars + 4);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace tcwgrad
}  // namespace endo
