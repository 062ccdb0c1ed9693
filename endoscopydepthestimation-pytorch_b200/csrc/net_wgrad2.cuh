// Weight gradient of the 3x3 / 1x1 convolutions on tcgen05, TMA-fed (round 2).
//
//   dW[co][ci][ky][kx] = sum_{p} act[p][ci] * G[p - (ky-1, kx-1)][co],   act = relu(bn(x)) (or the upsampled map), zero outside
//
// Same GEMM as tcwgrad::dense_wgrad_bf16_kernel (pixels = K, bf16 MN-major operands, three resident accumulator sets,
// halo on the small gradient operand), different data movement.  The round-1 kernel staged the activations with register
// loads: load batch -> wait -> transform -> store, three exposed DRAM round trips per 8x32 tile (ncu r2: 53 % of all
// warp-stall samples on the first use of a load).  Deeper register prefetch does not help on this machine: in-flight loads are
// tracked by six scoreboard counters per warp, and refilling a register buffer waits for the counter it shares with the batch
// still in flight (measured on the data-gradient kernel: +39 %).  Here the UNTRANSFORMED activation tile travels by TMA:
//
//   warp 17, one lane : cp.async.bulk.tensor box (64 channels, 32 x 8 pixels) of the NHWC level buffer -> raw ring (2 x 64 KB),
//                       up to two tiles ahead of the consumers, out-of-image pixels zero-filled by the hardware;
//   warps 0-15        : raw fp32 (shared) -> BatchNorm + ReLU -> bf16 -> operand planes (shared), then the gradient halo tile
//                       (registers, prefetched ONE tile ahead: a single batch in flight) -> three kx-shifted planes;
//   warp 16           : 51 MMAs per tile (17 K-steps x 3 vertical taps), accumulators resident in TMEM across all tiles.
//
// The operand stage is single (the raw ring took its place in shared memory): transform and MMAs of consecutive tiles
// alternate, ~4 k cycles per tile together, against ~10 k for the load-latency-bound round-1 loop.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"
#include "net_tc.cuh"

namespace endo {
namespace tcwgrad2 {

using tcwgrad::Args; using tcwgrad::PITCH; using tcwgrad::TR; using tcwgrad::TW; using tcwgrad::KPX; using tcwgrad::A_ROWS;
using tcwgrad::PLANE_BYTES; using tcwgrad::MCH; using tcwgrad::NB; using tcwgrad::pack_bf16;

constexpr int RAW_BYTES = TR * TW * MCH * 4;             // 65,536: [8][32][64] fp32
constexpr int NRAW = 2;
constexpr int A_BYTES = (MCH / 8) * PLANE_BYTES;         // 44,160
constexpr int G_BYTES = 6 * PLANE_BYTES;                 // 33,120
constexpr int OP_OFF = NRAW * RAW_BYTES;                 // 131,072
constexpr int PAD_BYTES = 2 * PLANE_BYTES;               // the M = 128 read of the A operand runs 16 planes far
constexpr int KTAB_OFF = OP_OFF + A_BYTES + G_BYTES + PAD_BYTES;
constexpr int KTAB_BYTES = 2 * 8 * 144;
constexpr int BAR_OFF = KTAB_OFF + KTAB_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256;
constexpr int NPROD = 512;
constexpr int NTHREADS = NPROD + 64;                     // + MMA warp + TMA warp

__global__ void __launch_bounds__(NTHREADS, 1)
dense_wgrad_tma_kernel(const Args A, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_s = smem + OP_OFF;
    unsigned char* g_s = a_s + A_BYTES;
    float* ktab = reinterpret_cast<float*>(smem + KTAB_OFF);                    // [G <= 2][8 groups][8 x float4 + pad]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);               // raw_full[2], raw_empty[2], op_full, op_empty, accum
    uint64_t* raw_full = bars; uint64_t* raw_empty = bars + 2; uint64_t* op_full = bars + 4; uint64_t* op_empty = bars + 5;
    uint64_t* accum = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.y * MCH;
    const int tiles_x = (A.W + TW - 1) / TW, tiles_y = (A.H + TR - 1) / TR;
    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    const int ntiles = t_end - t_begin;
    const int sh = A.up ? 1 : 0;

    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    if (tid < 2 * MCH) {                                     // coefficient table (zeros for TransitionUp: no BatchNorm in front)
        const int gi = tid / MCH, cl = tid % MCH, ch = ci0 + cl;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < A.G && ch < A.Cin && !A.up) e = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)gi * A.Cin + ch) * 4));
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(ktab) + (gi * 8 + (cl >> 3)) * 144 + (cl & 7) * 16) = e;
    }
    if (tid == 0) {
        tc::mbar_init(raw_full + 0, 1); tc::mbar_init(raw_full + 1, 1);
        tc::mbar_init(raw_empty + 0, NPROD); tc::mbar_init(raw_empty + 1, NPROD);
        tc::mbar_init(op_full, NPROD); tc::mbar_init(op_empty, 1); tc::mbar_init(accum, 1);
        tc::fence_mbar_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
        b = t / (tiles_x * tiles_y);
        const int rem = t - b * (tiles_x * tiles_y);
        const int tx = rem / tiles_y, ty = rem - tx * tiles_y;      // column-major: consecutive tiles of a CTA are vertical neighbours
        y0 = ty * TR; x0 = tx * TW;                                 // (the gradient halo rows they share hit the L2)
    };

    if (warp < 16) {
        // The operand stage is cleared ONCE: every tile writes the same rows (activation planes: the 8x32 interior; gradient
        // planes: rows kx .. 339 + kx of plane kx), every other row -- pad columns, margins of the shifted planes -- stays zero.
        for (int i = tid; i < (A_BYTES + G_BYTES + PAD_BYTES) / 16; i += NPROD) reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("bar.sync 1, 512;" ::: "memory");

        // ---- gradient halo tile, dense / upsampled modes: (8 + 2) x 34 pixels x two 8-channel halves = 680 items; a thread owns
        //      item tid and, if tid < 168, item 512 + tid.  The loads of tile it + 1 are issued after tile it has been written.
        const int q_a = tid >= A_ROWS ? tid - A_ROWS : tid, hf_a = tid >= A_ROWS ? 1 : 0;
        const int q_b = (NPROD + tid) - A_ROWS, hf_b = 1;                       // second item (valid when tid < 2 * A_ROWS - NPROD)
        const bool has_b = tid < 2 * A_ROWS - NPROD;
        float4 ga[2], xa4[2], gb[2], xb4[2];
        bool ok_a = false, ok_b = false;
        auto g_issue = [&](int t) {
            int b, y0, x0;
            tile_origin(t, b, y0, x0);
            const size_t img = (size_t)b * A.H * A.W;
            {
                const int r = q_a / PITCH, cc = q_a - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                ok_a = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                const size_t oo = (img + (size_t)y * A.W + x) * A.C + A.out_off + hf_a * 8;
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    ga[h4] = make_float4(0.f, 0.f, 0.f, 0.f); xa4[h4] = ga[h4];
                    if (ok_a && hf_a * 8 + h4 * 4 < A.Cout) {
                        ga[h4] = __ldg(reinterpret_cast<const float4*>(A.g + oo + h4 * 4));
                        xa4[h4] = __ldg(reinterpret_cast<const float4*>(A.x + oo + h4 * 4));
                    }
                }
            }
            if (has_b) {
                const int r = q_b / PITCH, cc = q_b - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                ok_b = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                const size_t oo = (img + (size_t)y * A.W + x) * A.C + A.out_off + hf_b * 8;
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    gb[h4] = make_float4(0.f, 0.f, 0.f, 0.f); xb4[h4] = gb[h4];
                    if (ok_b && hf_b * 8 + h4 * 4 < A.Cout) {
                        gb[h4] = __ldg(reinterpret_cast<const float4*>(A.g + oo + h4 * 4));
                        xb4[h4] = __ldg(reinterpret_cast<const float4*>(A.x + oo + h4 * 4));
                    }
                }
            }
        };
        auto g_store = [&](int g, int q, int hf, bool ok, const float4 (&gq)[2], const float4 (&xq)[2]) {
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
        };
        if (!A.one && ntiles > 0) g_issue(t_begin);

        const int grp = tid & 7;
        const int ch = ci0 + grp * 8;
        const bool ch_ok = ch < A.Cin;                           // Cin is a multiple of 4: a group may be half valid
        const bool hi_ok = ch + 4 < A.Cin;
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1;
            int b, y0, x0;
            tile_origin(t_begin + it, b, y0, x0);
            const int g = b / (A.B / A.G);
            tc::mbar_wait(raw_full + s, (it >> 1) & 1);                         // the TMA box of this tile has landed
            if (it >= 1) tc::mbar_wait(op_empty, (it - 1) & 1);                 // the MMAs of the previous tile are done with the planes
            // ---- activations: raw fp32 (shared) -> BN + ReLU -> bf16 planes.  Item = (interior pixel, 8-channel group).
            {
                const float4* kt = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(ktab) + (g * 8 + grp) * 144);
                const float4 k0 = kt[0], k1 = kt[1], k2 = kt[2], k3 = kt[3], k4 = kt[4], k5 = kt[5], k6 = kt[6], k7 = kt[7];
                const unsigned char* raw = smem + s * RAW_BYTES + grp * 32;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ip = (tid >> 3) + 64 * j;              // interior pixel 0 .. 255
                    const int ry = ip >> 5, rx = ip & 31;
                    const int px = (1 + ry) * PITCH + 1 + rx;        // row of the pitch-34 plane
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (ch_ok && y0 + ry < A.H && x0 + rx < A.W) {   // (TMA zero-fills outside the image, but relu(bn(0)) != 0)
                        const int sp = sh ? ((ry >> 1) * (TW >> 1) + (rx >> 1)) : ip;
                        const float4 a0 = *reinterpret_cast<const float4*>(raw + (size_t)sp * (MCH * 4));
                        const float4 a1 = *reinterpret_cast<const float4*>(raw + (size_t)sp * (MCH * 4) + 16);
                        float v0 = a0.x, v1 = a0.y, v2 = a0.z, v3 = a0.w, v4 = a1.x, v5 = a1.y, v6 = a1.z, v7 = a1.w;
                        if (!A.up) {
                            v0 = fmaxf(fmaf(k0.x, a0.x - k0.z, k0.y), 0.f); v1 = fmaxf(fmaf(k1.x, a0.y - k1.z, k1.y), 0.f);
                            v2 = fmaxf(fmaf(k2.x, a0.z - k2.z, k2.y), 0.f); v3 = fmaxf(fmaf(k3.x, a0.w - k3.z, k3.y), 0.f);
                            v4 = fmaxf(fmaf(k4.x, a1.x - k4.z, k4.y), 0.f); v5 = fmaxf(fmaf(k5.x, a1.y - k5.z, k5.y), 0.f);
                            v6 = fmaxf(fmaf(k6.x, a1.z - k6.z, k6.y), 0.f); v7 = fmaxf(fmaf(k7.x, a1.w - k7.z, k7.y), 0.f);
                        }
                        if (!hi_ok) v4 = v5 = v6 = v7 = 0.f;          // channels past Cin (the box may cover a neighbouring region)
                        o = make_uint4(pack_bf16(v0, v1), pack_bf16(v2, v3), pack_bf16(v4, v5), pack_bf16(v6, v7));
                    }
                    *reinterpret_cast<uint4*>(a_s + grp * PLANE_BYTES + (size_t)px * 16) = o;
                }
            }
            tc::mbar_arrive(raw_empty + s);                                     // this thread has read its part of the raw tile
            // ---- output gradient
            if (A.one) {
                // 1x1 mode (TransitionDown): interior only, 48 channels in three rounds of 16: the max-pool-routed gradient of the
                // NEXT level (argmax word + g + x); plane (sub, half) row q + 1
                const int gpix = tid & 255, half = tid >> 8;
                const int r = 1 + (gpix >> 5), cc = 1 + (gpix & 31);
                const int y = y0 + r - 1, x = x0 + cc - 1;
                const bool ok = (y < A.H) && (x < A.W);
                const unsigned pos = (unsigned)(((y & 1) << 1) | (x & 1));
                const size_t pp = ok ? ((size_t)(b * A.cH + (y >> 1)) * A.cW + (x >> 1)) : 0;
                const int q = r * PITCH + cc;
                unsigned am[3][2];
                float4 gq[3][2], xq[3][2];
#pragma unroll
                for (int sub = 0; sub < 3; ++sub) {                              // all loads of the three rounds first
                    const int cbase = A.out_off + sub * 16 + half * 8;
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        am[sub][h4] = 0xffffffffu; gq[sub][h4] = make_float4(0.f, 0.f, 0.f, 0.f); xq[sub][h4] = gq[sub][h4];
                        if (ok && cbase + h4 * 4 < A.Cout) {
                            am[sub][h4] = __ldg(reinterpret_cast<const unsigned*>(A.argmax + pp * A.Cout + cbase + h4 * 4));
                            gq[sub][h4] = __ldg(reinterpret_cast<const float4*>(A.gc + pp * A.cC + A.c_off + cbase + h4 * 4));
                            xq[sub][h4] = __ldg(reinterpret_cast<const float4*>(A.xc + pp * A.cC + A.c_off + cbase + h4 * 4));
                        }
                    }
                }
#pragma unroll
                for (int sub = 0; sub < 3; ++sub) {
                    const int cbase = A.out_off + sub * 16 + half * 8;
                    float v[8];
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                        if (cbase + h4 * 4 < A.Cout) {
                            const float* abp = A.abc + ((size_t)g * A.cC + A.c_off + cbase + h4 * 4) * 2;
                            c0 = __ldg(reinterpret_cast<const float4*>(abp)); c1 = __ldg(reinterpret_cast<const float4*>(abp + 4));
                        }
                        const unsigned a_ = am[sub][h4];
                        v[h4 * 4 + 0] = ((a_ & 0xffu) == pos) ? gq[sub][h4].x + fmaf(c0.y, xq[sub][h4].x, c0.x) : 0.f;
                        v[h4 * 4 + 1] = (((a_ >> 8) & 0xffu) == pos) ? gq[sub][h4].y + fmaf(c0.w, xq[sub][h4].y, c0.z) : 0.f;
                        v[h4 * 4 + 2] = (((a_ >> 16) & 0xffu) == pos) ? gq[sub][h4].z + fmaf(c1.y, xq[sub][h4].z, c1.x) : 0.f;
                        v[h4 * 4 + 3] = ((a_ >> 24) == pos) ? gq[sub][h4].w + fmaf(c1.w, xq[sub][h4].w, c1.z) : 0.f;
                    }
                    const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    *reinterpret_cast<uint4*>(g_s + (sub * 2 + half) * PLANE_BYTES + (size_t)(q + 1) * 16) = ok ? o : make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                g_store(g, q_a, hf_a, ok_a, ga, xa4);
                if (has_b) g_store(g, q_b, hf_b, ok_b, gb, xb4);
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(op_full);
            if (!A.one && it + 1 < ntiles) g_issue(t_begin + it + 1);           // one batch in flight while the MMAs of this tile run
        }
        // ---- epilogue: D_ky[ci][kx*16 + co] -> atomicAdd into OIHW
        tc::mbar_wait(accum, 0);
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
    } else if (warp == 16) {
        // ---------------------------------------------------------------- MMA issuer: convergent; one elected lane issues
        const uint32_t tmem_b = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const uint32_t idesc = tc::instr_desc(tc::FMT_BF16, 128, NB, 1, 1);
        const uint32_t a_base = tc::smem_u32(a_s), g_base = tc::smem_u32(g_s);
        // MN-major: LBO = 8-pixel groups (128 B), SBO = 8-channel groups (planes)
        const uint64_t d_hi = tc::smem_desc(0, 128, PLANE_BYTES);
        const uint64_t a_d0 = d_hi | (uint64_t)(a_base >> 4), b_d0 = d_hi | (uint64_t)((g_base + 16u) >> 4);
        for (int it = 0; it < ntiles; ++it) {
            tc::mbar_wait(op_full, it & 1);
            tc::tc_fence_after();
            // consecutive K-steps rotate over three accumulator sets (9 independent chains): an MMA that accumulates into the
            // tile its predecessor wrote waits ~266 cycles for it
            if (A.one) {
#pragma unroll 1
                for (int k16 = 0; k16 < KPX / 16; ++k16)
                    tc::mma_f16_w(tmem_b + (k16 % 9) * NB, a_d0 + (uint64_t)(PITCH + k16 * 16), b_d0 + (uint64_t)(PITCH + k16 * 16), idesc,
                                (uint32_t)(it != 0 || k16 >= 9));
            } else {
#pragma unroll 1
                for (int k16 = 0; k16 < KPX / 16; ++k16) {
                    const int set = k16 % 3;
                    const uint64_t ad = a_d0 + (uint64_t)(PITCH + k16 * 16);
                    const uint32_t acc = (uint32_t)(it != 0 || k16 >= 3);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)          // act[p] pairs with G[p - (ky-1) rows]: B rows slide, A stays
                        tc::mma_f16_w(tmem_b + (set * 3 + ky) * NB, ad, b_d0 + (uint64_t)(PITCH + k16 * 16 - (ky - 1) * PITCH), idesc, acc);
                }
            }
            tc::tc_commit_w(op_empty);
        }
        tc::tc_commit_w(accum);
    } else {
        // ---------------------------------------------------------------- TMA issuer (warp 17, lane 0): runs up to NRAW tiles ahead
        if (lane == 0) {
            tma::prefetch_map(&xmap);
            const uint32_t box_bytes = (uint32_t)(MCH * (TW >> sh) * (TR >> sh) * 4);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it & 1;
                if (it >= NRAW) tc::mbar_wait(raw_empty + s, ((it >> 1) - 1) & 1);
                int b, y0, x0;
                tile_origin(t_begin + it, b, y0, x0);
                tc::mbar_expect_tx(raw_full + s, box_bytes);
                tma::load_4d(smem + s * RAW_BYTES, &xmap, A.in_off + ci0, x0 >> sh, y0 >> sh, b, raw_full + s);
                tc::mbar_arrive(raw_full + s);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace tcwgrad2
}  // namespace endo
