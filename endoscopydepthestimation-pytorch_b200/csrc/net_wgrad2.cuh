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
//   warp 17, one lane : cp.async.bulk.tensor box (64 channels, 16 x 8 pixels) of the NHWC level buffer -> raw ring (2 x 32 KB),
//                       up to two tiles ahead of the consumers, out-of-image pixels zero-filled by the hardware;
//   warps 0-15        : raw fp32 (shared) -> BatchNorm + ReLU -> bf16 -> operand planes (shared); the (8+2) x 18 gradient halo
//                       tile (g and x, 12 channels) arrives by cp.async one tile ahead in a ring of thread-private slots and
//                       is corrected / packed into three kx-shifted planes;
//   warp 16           : 27 MMAs per tile (9 K-steps x 3 vertical taps), accumulators resident in TMEM across all tiles.
//
// Everything is double-buffered (raw ring, operand stages, gradient ring): TMA, transform and MMAs of consecutive tiles
// overlap.  History (clock64 traces, r2): with 8x32 tiles, a single operand stage and the gradient rows prefetched in
// REGISTERS across the loop edge, 55 % of the tile period was the loop top waiting for those registers.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"
#include "net_tc.cuh"

namespace endo {
namespace tcwgrad2 {

// trace slots: [0] = tiles, [1] = start; producers (thread 0) 16 + 8 it + {0 loop top, 1 raw landed, 2 planes free, 3 activations
// written, 4 gradient written / arrived}; MMA warp 16 + 8 it + {5 operands ready, 6 issued}; TMA thread 16 + 8 it + 7 = issued
#ifdef ENDO_TRACE_BUILD
#define WG_TRACE(slot) do { if ((A.dbg & 8) && blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 2048) g_tc_trace[(slot)] = clock64(); } while (0)
#else
#define WG_TRACE(slot) do { } while (0)
#endif

using tcwgrad::Args; using tcwgrad::MCH; using tcwgrad::NB; using tcwgrad::pack_bf16;

constexpr int TR = 8, TW = 16, PITCH = TW + 2;           // 8 x 16 interior pixels per tile, rows of the operand planes 18 pixels apart
constexpr int KPX = TR * PITCH;                          // 144 pixels = 9 K-steps of 16
constexpr int H_ROWS = (TR + 2) * PITCH;                 // 180 pixels of the gradient halo tile
constexpr int PLANE_BYTES = 185 * 16;                    // 2,960 (= 16 mod 128: the 8 channel groups of a pixel spread over all banks)
constexpr int RAW_BYTES = TR * TW * MCH * 4;             // 32,768: [8][16][64] fp32
constexpr int A_BYTES = (MCH / 8) * PLANE_BYTES;         // 23,680
constexpr int G_BYTES = 6 * PLANE_BYTES;                 // 17,760
constexpr int STAGE = A_BYTES + G_BYTES;                 // 41,440
constexpr int PAD_BYTES = 2 * PLANE_BYTES;               // the M = 128 read of the A operand runs 16 planes far (stage 1: past its end)
constexpr int GITEMS = 2 * H_ROWS;                       // (halo pixel, 8-channel half) items of the gradient tile
constexpr int KTAB_BYTES = 2 * 8 * 144;
constexpr int NPROD = 512;
constexpr int NTHREADS = NPROD + 64;                     // + MMA warp + TMA warp
// Shared-memory layout, raw-ring depth chosen per mode (the TMA latency, ~4.5 k cycles, spans more than one ~3.5 k-cycle tile
// period: with two raw slots the consumers waited ~1.1 k cycles per tile for the box, clock64 trace r2):
//   [raw ring: nraw x 32 KB][operand stages 2 x 41,440 + pad][coefficient table][barriers][gradient ring 2 x gstage]
//   dense layers (Cout <= 12): the second channel quad of the upper half never exists -> compact gradient slots, nraw = 3
//   TransitionUp passes / growth 16 (Cout = 16): full slots, nraw = 2;   1x1 mode: no gradient ring, nraw = 4
__host__ __device__ inline int raw_depth(int Cout, int one) { return one ? 4 : (Cout <= 12 ? 3 : 2); }
__host__ __device__ inline int g_hi_items(int Cout) { return Cout <= 12 ? H_ROWS : GITEMS; }       // items that own a second quad
__host__ __device__ inline int g_stage_bytes(int Cout, int one) { return one ? 0 : (2 * GITEMS + 2 * g_hi_items(Cout)) * 16; }
__host__ inline size_t smem_bytes(int Cout, int one) {
    return (size_t)raw_depth(Cout, one) * RAW_BYTES + 2 * STAGE + PAD_BYTES + KTAB_BYTES + 256 + 2 * (size_t)g_stage_bytes(Cout, one);
}

__global__ void __launch_bounds__(NTHREADS, 1)
dense_wgrad_tma_kernel(const Args A, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int NRAW = raw_depth(A.Cout, A.one);
    const int OP_OFF = NRAW * RAW_BYTES;
    float* ktab = reinterpret_cast<float*>(smem + OP_OFF + 2 * STAGE + PAD_BYTES);   // [G <= 2][8 groups][8 x float4 + pad]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OP_OFF + 2 * STAGE + PAD_BYTES + KTAB_BYTES);
    uint64_t* raw_full = bars; uint64_t* raw_empty = bars + 4; uint64_t* op_full = bars + 8; uint64_t* op_empty = bars + 10;
    uint64_t* accum = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    unsigned char* gring = smem + OP_OFF + 2 * STAGE + PAD_BYTES + KTAB_BYTES + 256;
    const int GRING_BYTES = g_stage_bytes(A.Cout, A.one);
    const int ghi = g_hi_items(A.Cout);                      // slot regions of a gradient stage: g lo [GITEMS], g hi [ghi], x lo, x hi

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ci0 = blockIdx.y * MCH;
    const int tiles_x = (A.W + TW - 1) / TW, tiles_y = (A.H + TR - 1) / TR;
    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    const int ntiles = t_end - t_begin;
    const int sh = A.up ? 1 : 0;
    pdl_trigger();
    if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { tc::mbar_init(raw_full + i, 1); tc::mbar_init(raw_empty + i, NPROD); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(op_full + i, NPROD); tc::mbar_init(op_empty + i, 1); }
        tc::mbar_init(accum, 1);
        tc::fence_mbar_init();
    }
    // Both operand stages are cleared ONCE: every tile writes the same rows (activation planes: the 8x16 interior; gradient
    // planes: rows kx .. 179 + kx of plane kx), every other row -- pad columns, margins of the shifted planes -- stays zero.
    if (warp < 16)
        for (int i = tid; i < (2 * STAGE + PAD_BYTES) / 16; i += NPROD) reinterpret_cast<uint4*>(smem + OP_OFF)[i] = make_uint4(0u, 0u, 0u, 0u);
    pdl_wait();                                              // on-chip set-up above; global memory from here on
    if (tid == 0) { WG_TRACE(1); if ((A.dbg & 8) && blockIdx.x == 0 && blockIdx.y == 0) g_tc_trace[0] = ntiles; }
    if (tid < 2 * MCH) {                                     // coefficient table (zeros for TransitionUp: no BatchNorm in front)
        const int gi = tid / MCH, cl = tid % MCH, ch = ci0 + cl;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi < A.G && ch < A.Cin && !A.up) e = __ldg(reinterpret_cast<const float4*>(A.coef + ((size_t)gi * A.Cin + ch) * 4));
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(ktab) + (gi * 8 + (cl >> 3)) * 144 + (cl & 7) * 16) = e;
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

        // ---- gradient halo tile, dense / upsampled modes: (8 + 2) x 18 pixels x two 8-channel halves = 360 items, one per thread
        //      (tid < 360).  g and x of the item travel by cp.async into a two-deep ring of thread-private slots, requested one
        //      tile AHEAD (commit groups, not scoreboards, track them: registers prefetched across the loop edge stalled the loop
        //      top for 55 % of the tile period, clock64 trace r2).
        const bool g_thread = tid < GITEMS && !A.one;
        const int gq = tid % H_ROWS, ghf = tid / H_ROWS;
        auto g_issue = [&](int it) {                                            // tile t_begin + it -> ring[it & 1]
            if (g_thread && it < ntiles) {
                int b, y0, x0;
                tile_origin(t_begin + it, b, y0, x0);
                const int r = gq / PITCH, cc = gq - r * PITCH;
                const int y = y0 + r - 1, x = x0 + cc - 1;
                const bool ok = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                const size_t oo = ok ? (((size_t)b * A.H + y) * A.W + x) * A.C + A.out_off + ghf * 8 : 0;
                unsigned char* slot = gring + (it & 1) * GRING_BYTES + (size_t)tid * 16;
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    if (ghf * 8 + h4 * 4 < A.Cout) {                             // (compact slots: the quad does not exist otherwise)
                        const uint32_t nb = ok ? 16u : 0u;                       // 0 bytes = zero fill
                        tc::cp_async16(slot + (size_t)(h4 ? GITEMS : 0) * 16, A.g + oo + (nb ? h4 * 4 : 0), nb);
                        tc::cp_async16(slot + (size_t)(GITEMS + ghi + (h4 ? GITEMS : 0)) * 16, A.x + oo + (nb ? h4 * 4 : 0), nb);
                    }
                }
            }
            tc::cp_async_commit();                                              // one group per tile, empty or not
        };
        g_issue(0);

        const int grp = tid & 7;
        const int ch = ci0 + grp * 8;
        const bool ch_ok = ch < A.Cin;                           // Cin is a multiple of 4: a group may be half valid
        const bool hi_ok = ch + 4 < A.Cin;
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1;
            const int rs = it % NRAW, rph = (it / NRAW) & 1;                    // raw-ring slot and its phase
            unsigned char* a_s = smem + OP_OFF + s * STAGE;
            unsigned char* g_s = a_s + A_BYTES;
            int b, y0, x0;
            tile_origin(t_begin + it, b, y0, x0);
            const int g = b / (A.B / A.G);
            if (tid == 0) WG_TRACE(16 + 8 * it + 0);
            g_issue(it + 1);                                                    // next tile's gradient rows: in flight from here
            tc::mbar_wait(raw_full + rs, rph);                                  // the TMA box of this tile has landed
            if (tid == 0) WG_TRACE(16 + 8 * it + 1);
            if (it >= 2) tc::mbar_wait(op_empty + s, ((it >> 1) - 1) & 1);      // the MMAs of tile it - 2 are done with this stage
            if (tid == 0) WG_TRACE(16 + 8 * it + 2);
            // ---- activations: raw fp32 (shared) -> BN + ReLU -> bf16 planes.  Item = (interior pixel, 8-channel group).
            {
                const float4* kt = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(ktab) + (g * 8 + grp) * 144);
                const float4 k0 = kt[0], k1 = kt[1], k2 = kt[2], k3 = kt[3], k4 = kt[4], k5 = kt[5], k6 = kt[6], k7 = kt[7];
                const unsigned char* raw = smem + rs * RAW_BYTES + grp * 32;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int ip = (tid >> 3) + 64 * j;              // interior pixel 0 .. 127
                    const int ry = ip >> 4, rx = ip & 15;
                    const int px = (1 + ry) * PITCH + 1 + rx;        // row of the pitch-18 plane
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
            if (tid == 0) WG_TRACE(16 + 8 * it + 3);
            tc::mbar_arrive(raw_empty + rs);                                    // this thread has read its part of the raw tile
            // ---- output gradient
            if (A.one) {
                // 1x1 mode (TransitionDown): interior only, 48 channels in three rounds of 16: the max-pool-routed gradient of the
                // NEXT level (argmax word + g + x); plane (sub, half) row q + 1.  128 pixels x 2 halves = threads 0 .. 255.
                if (tid < 256) {
                    const int gpix = tid & 127, half = tid >> 7;
                    const int r = 1 + (gpix >> 4), cc = 1 + (gpix & 15);
                    const int y = y0 + r - 1, x = x0 + cc - 1;
                    const bool ok = (y < A.H) && (x < A.W);
                    const unsigned pos = (unsigned)(((y & 1) << 1) | (x & 1));
                    const size_t pp = ok ? ((size_t)(b * A.cH + (y >> 1)) * A.cW + (x >> 1)) : 0;
                    const int q = r * PITCH + cc;
                    unsigned am[3][2];
                    float4 gv[3][2], xv[3][2];
#pragma unroll
                    for (int sub = 0; sub < 3; ++sub) {                              // all loads of the three rounds first
                        const int cbase = A.out_off + sub * 16 + half * 8;
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            am[sub][h4] = 0xffffffffu; gv[sub][h4] = make_float4(0.f, 0.f, 0.f, 0.f); xv[sub][h4] = gv[sub][h4];
                            if (ok && cbase + h4 * 4 < A.Cout) {
                                am[sub][h4] = __ldg(reinterpret_cast<const unsigned*>(A.argmax + pp * A.Cout + cbase + h4 * 4));
                                gv[sub][h4] = __ldg(reinterpret_cast<const float4*>(A.gc + pp * A.cC + A.c_off + cbase + h4 * 4));
                                xv[sub][h4] = __ldg(reinterpret_cast<const float4*>(A.xc + pp * A.cC + A.c_off + cbase + h4 * 4));
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
                            v[h4 * 4 + 0] = ((a_ & 0xffu) == pos) ? gv[sub][h4].x + fmaf(c0.y, xv[sub][h4].x, c0.x) : 0.f;
                            v[h4 * 4 + 1] = (((a_ >> 8) & 0xffu) == pos) ? gv[sub][h4].y + fmaf(c0.w, xv[sub][h4].y, c0.z) : 0.f;
                            v[h4 * 4 + 2] = (((a_ >> 16) & 0xffu) == pos) ? gv[sub][h4].z + fmaf(c1.y, xv[sub][h4].z, c1.x) : 0.f;
                            v[h4 * 4 + 3] = ((a_ >> 24) == pos) ? gv[sub][h4].w + fmaf(c1.w, xv[sub][h4].w, c1.z) : 0.f;
                        }
                        const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                        *reinterpret_cast<uint4*>(g_s + (sub * 2 + half) * PLANE_BYTES + (size_t)(q + 1) * 16) = ok ? o : make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else {
                tc::cp_async_wait<1>();                                         // this thread's copies of tile `it` have landed
                if (g_thread) {
                    const unsigned char* slot = gring + s * GRING_BYTES + (size_t)tid * 16;
                    float v[8];
                    const float* abp = A.ab + ((size_t)g * A.C + A.out_off + ghf * 8) * 2;
#pragma unroll
                    for (int h4 = 0; h4 < 2; ++h4) {
                        if (ghf * 8 + h4 * 4 < A.Cout) {
                            const float4 gv = *reinterpret_cast<const float4*>(slot + (size_t)(h4 ? GITEMS : 0) * 16);
                            const float4 xv = *reinterpret_cast<const float4*>(slot + (size_t)(GITEMS + ghi + (h4 ? GITEMS : 0)) * 16);
                            const float4 c0 = __ldg(reinterpret_cast<const float4*>(abp + h4 * 8));
                            const float4 c1 = __ldg(reinterpret_cast<const float4*>(abp + h4 * 8 + 4));
                            v[h4 * 4 + 0] = gv.x + fmaf(c0.y, xv.x, c0.x); v[h4 * 4 + 1] = gv.y + fmaf(c0.w, xv.y, c0.z);
                            v[h4 * 4 + 2] = gv.z + fmaf(c1.y, xv.z, c1.x); v[h4 * 4 + 3] = gv.w + fmaf(c1.w, xv.w, c1.z);
                        } else {
                            v[h4 * 4 + 0] = v[h4 * 4 + 1] = v[h4 * 4 + 2] = v[h4 * 4 + 3] = 0.f;
                        }
                    }
                    // a pixel outside the image: g and x were zero-filled, but the lazy BatchNorm term A_c is not zero
                    const int r = gq / PITCH, cc = gq - r * PITCH;
                    const int y = y0 + r - 1, x = x0 + cc - 1;
                    const bool ok = (y >= 0) && (y < A.H) && (x >= 0) && (x < A.W);
                    const uint4 o = ok ? make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]))
                                       : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        *reinterpret_cast<uint4*>(g_s + (kx * 2 + ghf) * PLANE_BYTES + (size_t)(gq + kx) * 16) = o;
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(op_full + s);
            if (tid == 0) WG_TRACE(16 + 8 * it + 4);
        }
        tc::cp_async_wait<0>();
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
        // MN-major: LBO = 8-pixel groups (128 B), SBO = 8-channel groups (planes)
        const uint64_t d_hi = tc::smem_desc(0, 128, PLANE_BYTES);
        for (int it = 0; it < ntiles; ++it) {
            const int s = it & 1;
            tc::mbar_wait(op_full + s, (it >> 1) & 1);
            tc::tc_fence_after();
            if (lane == 0) WG_TRACE(16 + 8 * it + 5);
            const uint32_t a_base = tc::smem_u32(smem + OP_OFF + s * STAGE), g_base = a_base + A_BYTES;
            const uint64_t a_d0 = d_hi | (uint64_t)(a_base >> 4), b_d0 = d_hi | (uint64_t)((g_base + 16u) >> 4);
            // consecutive K-steps rotate over three accumulator sets (9 independent chains): an MMA that accumulates into the
            // tile its predecessor wrote waits ~266 cycles for it
            if (A.one) {
#pragma unroll 1
                for (int k16 = 0; k16 < KPX / 16; ++k16)
                    tc::mma_f16_w(tmem_b + (k16 % 9) * NB, a_d0 + (uint64_t)(PITCH + k16 * 16), b_d0 + (uint64_t)(PITCH + k16 * 16), idesc,
                                (uint32_t)(it != 0));
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
            tc::tc_commit_w(op_empty + s);
            if (lane == 0) WG_TRACE(16 + 8 * it + 6);
        }
        tc::tc_commit_w(accum);
    } else {
        // ---------------------------------------------------------------- TMA issuer (warp 17, lane 0): runs up to NRAW tiles ahead
        if (lane == 0) {
            tma::prefetch_map(&xmap);
            const uint32_t box_bytes = (uint32_t)(MCH * (TW >> sh) * (TR >> sh) * 4);
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % NRAW;
                if (it >= NRAW) tc::mbar_wait(raw_empty + s, ((it / NRAW) - 1) & 1);
                int b, y0, x0;
                tile_origin(t_begin + it, b, y0, x0);
                tc::mbar_expect_tx(raw_full + s, box_bytes);
                tma::load_4d(smem + s * RAW_BYTES, &xmap, A.in_off + ci0, x0 >> sh, y0 >> sh, b, raw_full + s);
                tc::mbar_arrive(raw_full + s);
                WG_TRACE(16 + 8 * it + 7);
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
