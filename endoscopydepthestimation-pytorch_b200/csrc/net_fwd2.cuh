// DenseLayer forward (reference models.py:19-28: BN -> ReLU -> conv3x3 Cin -> 12|16), 3xTF32, PERSISTENT and TMA-fed (round 2).
//
// Same implicit GEMM as tcconv::dense_fwd_tf32_kernel (M = 128 linear pixels of the halo tile, N = 3 kx x 16 co, K = 8 input
// channels per chunk as the planes hi0 | hi1 | lo | x, vertical taps = descriptor start-address offsets, horizontal taps
// finished by the epilogue), different data movement and scheduling.  What the round-1 kernel lost (clock64 trace, r2): 28 k of
// every ~124 k cycles of a 32x32 tile in prologue + epilogue (nothing overlaps them with 1 CTA/SM), and ~2 k of every ~4.2 k
// cycles per channel chunk in the producers waiting for their single register-staged load batch.  Here:
//
//   * one CTA per SM walks over 32 x 16 tiles; the accumulators (5 M-blocks x 48 columns) are DOUBLE-BUFFERED in TMEM, so the
//     epilogue of tile k overlaps the channel loop of tile k + 1;
//   * warp 17 (one lane) streams the UNTRANSFORMED (8 channels, 34 x 18 pixels) boxes of the NHWC level buffer with
//     cp.async.bulk.tensor into a 4-deep raw ring -- zero-filled outside the image, up to four chunks (78 KB) in flight per SM,
//     across tile boundaries -- and the weight stage images travel by TMA bulk copy;
//   * warps 0-15 turn raw fp32 into the operand planes (BatchNorm + ReLU from a shared coefficient table, exact hi/lo split);
//   * warp 16 issues 30 MMAs per chunk; warps 18-21 (one per TMEM lane quadrant) drain finished accumulators: horizontal taps,
//     bias, per-channel statistics in registers across ALL tiles of the CTA (one fp64 atomic set per CTA and statistic group
//     instead of one per tile), and the 32 x 16 x 12 output block goes to the level buffer as ONE TMA tensor store
//     (round 1: 3,072 scattered 16-byte stores per tile, 768 bytes apart).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"
#include "net_tc.cuh"

namespace endo {
namespace tcfwd2 {

using tcconv::tf32_hi; using tcconv::bf16x2_rn;

constexpr int TH = 16, TW = 32, PITCH = 34;
constexpr int HALO_ROWS = PITCH * (TH + 2);              // 612 halo-tile pixels
constexpr int MBLK = 5;                                  // ceil(TH * PITCH / 128)
constexpr int PLANE_ROWS = 716;                          // >= 34 + 5*128 + 34 ; 716*16 % 128 == 64: the two hi planes a warp stores to (lane parity) fall in disjoint bank halves
constexpr int PLANE_BYTES = PLANE_ROWS * 16;             // 11,456
constexpr int A_STAGE = 4 * PLANE_BYTES;                 // 45,824: hi0 | hi1 | lo | x
constexpr int NB = 48;
constexpr int B_BLOCK = 2 * NB * 16;                     // 1,536
constexpr int B_STAGE = 6 * B_BLOCK;                     // 9,216 = one packed weight chunk (pack_w_fwd_all_kernel, mode 1)
constexpr int RAW_BYTES = HALO_ROWS * 32;                // 19,584 = 153 * 128: [18][34][8] fp32
constexpr int UP_W = TW / 2 + 2, UP_H = TH / 2 + 2;      // half-resolution box of the upsampling mode: 18 x 10 pixels
constexpr int RAW_BYTES_UP = UP_W * UP_H * 32;           // 5,760
constexpr int NRAW = 4;
constexpr int COEF_MAX = 384;
constexpr int OUT_MAXN = 16;
constexpr int NUNITS = MBLK * 4;
// shared-memory map
constexpr int RAW_OFF = 0;
constexpr int A_OFF = RAW_OFF + NRAW * RAW_BYTES;        // 78,336
constexpr int B_OFF = A_OFF + 2 * A_STAGE;               // 169,984
constexpr int OUT_OFF = B_OFF + 2 * B_STAGE;             // 188,416 (128-byte aligned: TMA store source)
constexpr int OUT_BYTES = TH * TW * OUT_MAXN * 4;        // 32,768 (N = 12: 24,576 used)
constexpr int COEF_OFF = OUT_OFF + OUT_BYTES;            // 221,184
constexpr int EDGE_OFF = COEF_OFF + COEF_MAX * 16;       // 227,328
constexpr int EDGE_BYTES = NUNITS * 2 * 16 * 4;          // 2,560
constexpr int RED_OFF = EDGE_OFF + EDGE_BYTES;           // 229,888
constexpr int RED_BYTES = 4 * 16 * 2 * 4;                // 512
constexpr int BAR_OFF = RED_OFF + RED_BYTES;             // 230,400
constexpr int SMEM_BYTES = BAR_OFF + 256;                // 230,656 <= 232,448
constexpr int NPROD = 512;
constexpr int NTHREADS = NPROD + 64 + 128;               // 16 transform warps + MMA warp + TMA warp + 4 epilogue warps

struct Args {
    const float* coef; const float* bias; const float* wpack; double* stats;
    int in_off, K, out_off, N, H, W, B, G, stats_C;
    int tiles_x, tiles_y, n_tiles;
    int up;                                              // 1: TransitionUp (models.py:70-80): the operand is the half-resolution buffer,
                                                         // nearest-upsampled x2, no BatchNorm / ReLU; the TMA box is (8, 18, 10) of it
    int dbg;                                             // ENDO_TC_DEBUG bit 16: clock64 trace of CTA 0 (tools/trace_fwd2.py)
};
// trace slots (first 3 tiles): [0] = chunks per tile, [1] = start, [2] = tiles; per running chunk j < 70: 16 + 8 j + {0 top, 1 raw landed,
// 2 stage free, 3 planes written (transform thread 0); 4 operands ready, 5 MMAs issued (MMA warp); 6 box issued (TMA thread)};
// per tile k < 8: 1900 + 4 k + {0 accumulators ready, 1 TMEM drained, 2 store issued} (epilogue thread 0)
// (compiled in only with -DENDO_TRACE_BUILD, `ENDO_BUILD_TRACE=1 python -m endo_b200.build --force`: the four predicated trace
// points of transform warp 0 cost 4.5 % of this kernel even when switched off at run time -- every other warp waits for it)
#ifdef ENDO_TRACE_BUILD
#define F2_TRACE(slot) do { if ((A.dbg & 16) && blockIdx.x == 0) g_tc_trace[(slot)] = clock64(); } while (0)
#else
#define F2_TRACE(slot) do { } while (0)
#endif

// NCHW images (<= 8 channels) -> NHWC fp32 with 8 channels (zeros past Cimg): the operand of the first convolution
// (models.py:111-113) in the layout the TMA boxes of the kernel below want
__global__ void __launch_bounds__(256)
nchw_to_nhwc8_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int Cimg, long long hw) {
    pdl_enter();
    const long long total = (long long)B * hw;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long b = i / hw, p = i - b * hw;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < Cimg ? __ldg(img + ((size_t)b * Cimg + c) * hw + p) : 0.f;
        float4* o = reinterpret_cast<float4*>(out + (size_t)i * 8);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
dense_fwd_x3_persistent_kernel(const Args A, const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* out_s = reinterpret_cast<float*>(smem + OUT_OFF);
    float* coef_s = reinterpret_cast<float*>(smem + COEF_OFF);
    float* edge = reinterpret_cast<float*>(smem + EDGE_OFF);
    float* red = reinterpret_cast<float*>(smem + RED_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
    uint64_t* raw_full = bars;            // [4]  count 1 + transaction bytes
    uint64_t* raw_empty = bars + 4;       // [4]  count NPROD
    uint64_t* op_full = bars + 8;         // [2]  count NPROD + transaction bytes (weights)
    uint64_t* op_empty = bars + 10;       // [2]  tcgen05.commit
    uint64_t* acc_full = bars + 12;       // [2]  tcgen05.commit
    uint64_t* acc_empty = bars + 14;      // [2]  count 128 (epilogue threads)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    __shared__ float s_bias[16];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool bn = !A.up && A.coef != nullptr;              // coef == nullptr: the operand is used as it is (first convolution)
    const int nchunks = (A.K + 7) >> 3;
    const int my_tiles = ((int)blockIdx.x < A.n_tiles) ? (A.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int per_img = A.tiles_x * A.tiles_y;
    const int per_group = A.B / A.G;

    pdl_trigger();
    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) { tc::mbar_init(raw_full + i, 1); tc::mbar_init(raw_empty + i, NPROD); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(op_full + i, NPROD); tc::mbar_init(op_empty + i, 1);
            tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 128);
        }
        tc::fence_mbar_init();
    }
    if (warp == 17 && lane == 0) { tma::prefetch_map(&in_map); tma::prefetch_map(&out_map); }
    pdl_wait();                                              // on-chip set-up above; global memory from here on
    if (tid == 0 && (A.dbg & 16) && blockIdx.x == 0) { g_tc_trace[0] = nchunks; g_tc_trace[1] = clock64(); g_tc_trace[2] = my_tiles; }
    if (tid < 16) s_bias[tid] = (tid < A.N) ? __ldg(A.bias + tid) : 0.f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto tile_origin = [&](int k, int& b, int& y0, int& x0) {       // k-th tile of this CTA
        const int t = (int)blockIdx.x + k * (int)gridDim.x;
        b = t / per_img;
        const int rem = t - b * per_img;
        const int ty = rem / A.tiles_x, tx = rem - ty * A.tiles_x;
        y0 = ty * TH; x0 = tx * TW;
    };

    if (warp < 16) {
        // ======================================================================== transform warps
        const int quad = tid & 1;                                   // 4-channel group inside the 8-channel chunk
        int cur_g = -1;
        int j = 0;                                                  // running chunk number of this CTA (all tiles)
        for (int k = 0; k < my_tiles; ++k) {
            int b, y0, x0;
            tile_origin(k, b, y0, x0);
            const int g = b / per_group;
            if (g != cur_g && bn) {
                // (a, beta, mean, invstd) of every input channel of this statistic group.  The table is only read by these 512
                // threads, between their own barriers: safe to rewrite when the group changes (tiles are visited in image order).
                asm volatile("bar.sync 1, 512;" ::: "memory");
                for (int i = tid; i < A.K; i += NPROD)
                    *reinterpret_cast<float4*>(coef_s + i * 4) = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)g * A.K + i) * 4));
                asm volatile("bar.sync 1, 512;" ::: "memory");
                cur_g = g;
            }
            unsigned pixok = 0u;                                    // which of this thread's 3 halo pixels lie inside the image
#pragma unroll
            for (int r3 = 0; r3 < 3; ++r3) {
                const int px = (tid + NPROD * r3) >> 1;
                const int r = px / PITCH, cc = px - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                if (px < HALO_ROWS && y >= 0 && y < A.H && x >= 0 && x < A.W) pixok |= 1u << r3;
            }
            for (int c = 0; c < nchunks; ++c, ++j) {
                const int rs = j & (NRAW - 1), s = j & 1;
                const int ch = c * 8 + quad * 4;
                const bool ch_ok = ch < A.K;
                float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0, k3 = k0;   // (a, beta, mean, invstd) x 4 channels
                if (ch_ok && bn) {
                    const float* cf = coef_s + ch * 4;
                    k0 = *reinterpret_cast<const float4*>(cf); k1 = *reinterpret_cast<const float4*>(cf + 4);
                    k2 = *reinterpret_cast<const float4*>(cf + 8); k3 = *reinterpret_cast<const float4*>(cf + 12);
                }
                if (tid == 0 && j < 70) F2_TRACE(16 + 8 * j + 0);
                tc::mbar_wait(raw_full + rs, (j >> 2) & 1);                        // the box of this chunk has landed
                if (tid == 0 && j < 70) F2_TRACE(16 + 8 * j + 1);
                if (j >= 2) tc::mbar_wait(op_empty + s, ((j >> 1) - 1) & 1);       // the MMAs of chunk j - 2 are done with the stage
                if (tid == 0 && j < 70) F2_TRACE(16 + 8 * j + 2);
                if (tid == 0) {                                                    // weights of this chunk: TMA bulk copy of the stage image
                    tc::mbar_expect_tx(op_full + s, (uint32_t)B_STAGE);
                    tc::bulk_g2s(smem + B_OFF + s * B_STAGE, A.wpack + (size_t)c * (B_STAGE / 4), (uint32_t)B_STAGE, op_full + s);
                }
                const unsigned char* raw = smem + RAW_OFF + rs * RAW_BYTES;
                unsigned char* a_s = smem + A_OFF + s * A_STAGE;
                // BRANCH-FREE item loop (the three items of a thread interleave in the instruction stream: the transform, not the MMA
                // issue or the TMA stream, bounded this kernel -- clock64 trace r2): loads are unconditional (clamped index), invalid
                // items become zeros through a mask, only the stores of the partial third round are predicated.  The tf32 "hi" part is
                // round-to-nearest-ties-away done on the integer pipe ((bits + 0x1000) & ~0x1fff == cvt.rna.tf32.f32 for finite
                // normal values), and the bf16 copy of x that multiplies w_lo (2^-12 of the product) is the upper half of hi.
                float4 v3[3];
#pragma unroll
                for (int r3 = 0; r3 < 3; ++r3) {
                    const int i = min(tid + NPROD * r3, 2 * HALO_ROWS - 1);
                    if (A.up) {                                          // fine pixel (r, cc) of the halo tile <- half-resolution box pixel
                        const int px = i >> 1, r = px / PITCH, cc = px - r * PITCH;
                        const int sp = (((r - 1) >> 1) + 1) * UP_W + ((cc - 1) >> 1) + 1;
                        v3[r3] = *reinterpret_cast<const float4*>(raw + (size_t)sp * 32 + quad * 16);
                    } else {
                        v3[r3] = *reinterpret_cast<const float4*>(raw + (size_t)i * 16);
                    }
                }
#pragma unroll
                for (int r3 = 0; r3 < 3; ++r3) {
                    const int i = tid + NPROD * r3;
                    const int px = i >> 1;
                    float4 v = v3[r3];
                    if (bn) {
                        v.x = fmaxf(fmaf(k0.x, v.x - k0.z, k0.y), 0.f); v.y = fmaxf(fmaf(k1.x, v.y - k1.z, k1.y), 0.f);
                        v.z = fmaxf(fmaf(k2.x, v.z - k2.z, k2.y), 0.f); v.w = fmaxf(fmaf(k3.x, v.w - k3.z, k3.y), 0.f);
                    }
                    const bool live = (pixok & (1u << r3)) && ch_ok;
                    v.x = live ? v.x : 0.f; v.y = live ? v.y : 0.f; v.z = live ? v.z : 0.f; v.w = live ? v.w : 0.f;
                    const uint32_t hx = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u, hy = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u;
                    const uint32_t hz = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u, hw = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u;
                    const float4 hi = make_float4(__uint_as_float(hx), __uint_as_float(hy), __uint_as_float(hz), __uint_as_float(hw));
                    const uint2 lo = make_uint2(bf16x2_rn(v.x - hi.x, v.y - hi.y), bf16x2_rn(v.z - hi.z, v.w - hi.w));
                    const uint2 xb = make_uint2(__byte_perm(hx, hy, 0x7632), __byte_perm(hz, hw, 0x7632));   // bf16(x) ~ upper halves of hi
                    if (r3 < 2 || px < HALO_ROWS) {
                        *reinterpret_cast<float4*>(a_s + quad * PLANE_BYTES + (size_t)px * 16) = hi;
                        *reinterpret_cast<uint2*>(a_s + 2 * PLANE_BYTES + (size_t)px * 16 + quad * 8) = lo;
                        *reinterpret_cast<uint2*>(a_s + 3 * PLANE_BYTES + (size_t)px * 16 + quad * 8) = xb;
                    }
                }
                tc::mbar_arrive(raw_empty + rs);
                tc::fence_proxy_async();
                tc::mbar_arrive(op_full + s);
                if (tid == 0 && j < 70) F2_TRACE(16 + 8 * j + 3);
            }
        }
    } else if (warp == 16) {
        // ======================================================================== MMA issuer: convergent; one elected lane issues
        const uint32_t tmem_b = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t idesc = tc::instr_desc(tc::FMT_TF32, 128, NB), idesc_b = tc::instr_desc(tc::FMT_BF16, 128, NB);
        const uint64_t a_hi = tc::smem_desc(0, PLANE_BYTES, 128), b_hi = tc::smem_desc(0, NB * 16, 128);
        int j = 0;
        for (int k = 0; k < my_tiles; ++k) {
            const int buf = k & 1;
            if (k >= 2) tc::mbar_wait(acc_empty + buf, ((k >> 1) - 1) & 1);        // the epilogue of tile k - 2 has drained the buffer
            tc::tc_fence_after();
            const uint32_t d0 = tmem_b + (uint32_t)(buf * MBLK * NB);
            for (int c = 0; c < nchunks; ++c, ++j) {
                const int s = j & 1;
                tc::mbar_wait(op_full + s, (j >> 1) & 1);
                tc::tc_fence_after();
                if (lane == 0 && j < 70) F2_TRACE(16 + 8 * j + 4);
                const uint32_t a_base = tc::smem_u32(smem + A_OFF + s * A_STAGE);
                const uint32_t b_base = tc::smem_u32(smem + B_OFF + s * B_STAGE);
#pragma unroll 1
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t row0 = (uint32_t)(PITCH + (ky - 1) * PITCH) * 16u;
                    const uint64_t ahi = a_hi | (uint64_t)((a_base + row0) >> 4);
                    const uint64_t alo = a_hi | (uint64_t)((a_base + 2u * PLANE_BYTES + row0) >> 4);
                    const uint64_t bhi = b_hi | (uint64_t)((b_base + (uint32_t)(ky * 2 + 0) * B_BLOCK) >> 4);
                    const uint64_t blo = b_hi | (uint64_t)((b_base + (uint32_t)(ky * 2 + 1) * B_BLOCK) >> 4);
                    const uint32_t acc = (uint32_t)((c | ky) != 0);
                    // both cross terms in ONE kind::f16 MMA of K = 16: [lo ; x] (planes 2, 3) x [w ; w - hi]; then hi * hi (tf32)
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb) tc::mma_f16_w(d0 + mb * NB, alo + (uint64_t)(mb * 128), blo, idesc_b, acc);
#pragma unroll
                    for (int mb = 0; mb < MBLK; ++mb) tc::mma_tf32_w(d0 + mb * NB, ahi + (uint64_t)(mb * 128), bhi, idesc, 1u);
                }
                tc::tc_commit_w(op_empty + s);
                if (lane == 0 && j < 70) F2_TRACE(16 + 8 * j + 5);
            }
            tc::tc_commit_w(acc_full + buf);
        }
    } else if (warp == 17) {
        // ======================================================================== TMA issuer: up to NRAW chunks ahead, across tiles
        if (lane == 0) {
            tma::prefetch_map(&in_map);
            int j = 0;
            for (int k = 0; k < my_tiles; ++k) {
                int b, y0, x0;
                tile_origin(k, b, y0, x0);
                for (int c = 0; c < nchunks; ++c, ++j) {
                    const int rs = j & (NRAW - 1);
                    if (j >= NRAW) tc::mbar_wait(raw_empty + rs, ((j >> 2) - 1) & 1);
                    tc::mbar_expect_tx(raw_full + rs, (uint32_t)(A.up ? RAW_BYTES_UP : RAW_BYTES));
                    if (A.up) tma::load_4d(smem + RAW_OFF + rs * RAW_BYTES, &in_map, A.in_off + c * 8, (x0 >> 1) - 1, (y0 >> 1) - 1, b, raw_full + rs);
                    else tma::load_4d(smem + RAW_OFF + rs * RAW_BYTES, &in_map, A.in_off + c * 8, x0 - 1, y0 - 1, b, raw_full + rs);
                    tc::mbar_arrive(raw_full + rs);
                    if (j < 70) F2_TRACE(16 + 8 * j + 6);
                }
            }
        }
    } else {
        // ======================================================================== epilogue warps 18-21: TMEM lane quadrant q
        const int q = warp & 3, et = tid - 18 * 32;                   // a warp reads the TMEM lanes 32 (warp % 4) .. + 31; et = 0 .. 127
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float s1[16], s2[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) { s1[jj] = 0.f; s2[jj] = 0.f; }
        int cur_g = -1;
        auto flush_stats = [&](int g) {                              // warp tree -> shared -> one fp64 atomic per channel
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                float a = s1[jj], c2 = s2[jj];
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o2);
                    c2 += __shfl_xor_sync(0xffffffffu, c2, o2);
                }
                if (lane == 0) { red[(q * 16 + jj) * 2] = a; red[(q * 16 + jj) * 2 + 1] = c2; }
                s1[jj] = 0.f; s2[jj] = 0.f;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et < 2 * A.N) {
                const int jj = et >> 1, which = et & 1;
                double sum = 0.0;
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) sum += (double)red[(wq * 16 + jj) * 2 + which];
                atomicAdd(A.stats + ((size_t)g * A.stats_C + A.out_off + jj) * 2 + which, sum);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
        };
        for (int k = 0; k < my_tiles; ++k) {
            const int buf = k & 1;
            int b, y0, x0;
            tile_origin(k, b, y0, x0);
            const int g = b / per_group;
            if (g != cur_g) { if (cur_g >= 0) flush_stats(cur_g); cur_g = g; }
            const uint32_t d0 = tmem + lane_base + (uint32_t)(buf * MBLK * NB);
            tc::mbar_wait(acc_full + buf, (k >> 1) & 1);
            tc::tc_fence_after();
            if (et == 0 && k < 8) F2_TRACE(1900 + 4 * k + 0);
            // the staging block of the previous tile must have been read by its TMA store before it is overwritten
            if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            // pass 1: publish the values the neighbouring 32-lane units need
#pragma unroll 1
            for (int mb = 0; mb < MBLK; ++mb) {
                const int u = mb * 4 + q;
                uint32_t r0[16], r2[16];
                tc::tmem_ld16_issue(d0 + mb * NB + 0, r0);
                tc::tmem_ld16_issue(d0 + mb * NB + 32, r2);
                tc::tmem_ld_wait();
                if (lane == 31) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) edge[(u * 2 + 1) * 16 + jj] = __uint_as_float(r0[jj]);    // kx = 0 part of my last pixel
                }
                if (lane == 0) {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) edge[(u * 2 + 0) * 16 + jj] = __uint_as_float(r2[jj]);    // kx = 2 part of my first pixel
                }
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");               // edges visible; staging block free (thread 0 waited above)
#pragma unroll 1
            for (int mb = 0; mb < MBLK; ++mb) {
                const int u = mb * 4 + q;
                const int L = PITCH + mb * 128 + q * 32 + lane;           // linear index in the halo tile
                const int r = L / PITCH, cc = L - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                const bool in_tile = (r >= 1) && (r <= TH) && (cc >= 1) && (cc <= TW);
                const bool ok = in_tile && (y < A.H) && (x < A.W);
                const float* eL = edge + ((u > 0 ? u - 1 : 0) * 2 + 1) * 16;
                const float* eR = edge + ((u < NUNITS - 1 ? u + 1 : u) * 2 + 0) * 16;
                const float keepL = (u > 0) ? 1.f : 0.f, keepR = (u < NUNITS - 1) ? 1.f : 0.f;
                float* op = out_s + ((r - 1) * TW + (cc - 1)) * A.N;
                // two halves of 8 channels (register budget: 704 threads per CTA leave 88 registers per thread)
#pragma unroll
                for (int h8 = 0; h8 < 16; h8 += 8) {
                    float v0[8], v1[8], v2[8];
                    tc::tmem_ld8(d0 + mb * NB + 0 + h8, v0);
                    tc::tmem_ld8(d0 + mb * NB + 16 + h8, v1);
                    tc::tmem_ld8(d0 + mb * NB + 32 + h8, v2);
                    if (mb == MBLK - 1 && h8 == 8) {                       // last TMEM read of this tile: tile k + 2 may overwrite the buffer
                        tc::tc_fence_before();
                        tc::mbar_arrive(acc_empty + buf);
                        if (et == 0 && k < 8) F2_TRACE(1900 + 4 * k + 1);
                    }
                    float o[8];
#pragma unroll
                    for (int j4 = 0; j4 < 8; j4 += 4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(eL + h8 + j4), r4 = *reinterpret_cast<const float4*>(eR + h8 + j4);
                        const float lq[4] = {l4.x * keepL, l4.y * keepL, l4.z * keepL, l4.w * keepL};
                        const float rq[4] = {r4.x * keepR, r4.y * keepR, r4.z * keepR, r4.w * keepR};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int jj = j4 + e;
                            float left = __shfl_up_sync(0xffffffffu, v0[jj], 1);
                            float right = __shfl_down_sync(0xffffffffu, v2[jj], 1);
                            left = (lane == 0) ? lq[e] : left;
                            right = (lane == 31) ? rq[e] : right;
                            o[jj] = (left + v1[jj]) + right + s_bias[h8 + jj];
                        }
                    }
                    if (in_tile) {                                        // staging block [16][32][N]; the store clips at the image edge
#pragma unroll
                        for (int jj = 0; jj < 8; jj += 4)
                            if (h8 + jj < A.N) *reinterpret_cast<float4*>(op + h8 + jj) = make_float4(o[jj], o[jj + 1], o[jj + 2], o[jj + 3]);
                    }
                    if (ok) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            if (h8 + jj < A.N) { s1[h8 + jj] += o[jj]; s2[h8 + jj] += o[jj] * o[jj]; }
                    }
                }
            }
            tc::fence_proxy_async();                                      // staging writes -> visible to the TMA store
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (et == 0) {
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(&out_map),
                             "r"(A.out_off), "r"(x0), "r"(y0), "r"(b), "r"(tc::smem_u32(out_s))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (k < 8) F2_TRACE(1900 + 4 * k + 2);
            }
        }
        if (cur_g >= 0) flush_stats(cur_g);
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // the last store has been written
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace tcfwd2
}  // namespace endo
