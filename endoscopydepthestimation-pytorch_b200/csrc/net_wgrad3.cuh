// DenseLayer weight gradient (reference models.py:19-28 under loss.backward(), train.py:325) as a pure TMA -> tcgen05 GEMM.
//
//   dW[co][ci][ky][kx] = sum_p act[p][ci] * G[p - (ky-1, kx-1)][co]        M = ci (blocks of 128), N = 3 kx x 16 co, K = pixels
//
// The two operands are bf16 by-products of the data-gradient kernel of the SAME layer, which has to evaluate them anyway:
// act = relu(bn(x)) [pixels][Cin padded to 8] (its epilogue computes it for the ReLU mask) and G = the output gradient with the
// lazy BatchNorm term [pixels][16] (what it stages as its own operand).  So nothing is transformed here -- the r2 ncu profile of the
// previous kernel (net_wgrad2.cuh: fp32 activations by TMA, BatchNorm + ReLU + bf16 conversion by 16 transform warps) showed it
// bound by the instruction stream of those warps at 0.26 of the HBM roofline -- and half the bytes are read (2 instead of 4 per
// activation).  Per 8 x 16-pixel tile, one thread issues
//     * ONE box (16 pixels x 8 channels, 8 rows, all Cin/8 channel groups) of act: shared memory receives the MN-major
//       SWIZZLE_NONE planes [group][pixel][8 channels] the MMA descriptor addresses, out-of-image pixels zero-filled;
//     * NINE boxes (16 pixels x 8 channels, 8 rows, 2 groups) of G, shifted by (ky - 1, kx - 1): the halo sits on the small
//       operand (its re-reads hit the L2), and all nine taps are ONE N = 144 MMA per K-step and channel block.  (First version:
//       three 10-row boxes, vertical taps as descriptor offsets, N = 48 -> 3x the MMAs; the single issuing thread, ~230 cycles of
//       loop overhead per MMA, bounded the kernel at 200 us per full-resolution layer whatever Cin.)
// one thread issues 8 K-steps x (Cin/128) MMAs into TMEM accumulators that stay resident over all tiles of the CTA (rotated over
// up to three sets: an MMA waits for the previous writer of its columns), and four warps add them into dW at the end.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"
#include "net_tc.cuh"

namespace endo {
namespace tcwgrad3 {

constexpr int TR = 8, TW = 16, KPX = TR * TW;         // 128 pixels per tile = 8 K-steps of 16
constexpr int NB = 144;                               // MMA N: 3 ky x 3 kx x 16 co
constexpr int A_PLANE = KPX * 16;                     // 2,048: [pixel][8 channels bf16]
constexpr int G_PLANE = A_PLANE;                      // a shifted copy of the gradient tile per tap, same pixel rows as the activations
constexpr int G_BYTES = 18 * G_PLANE;                 // 36,864: [ky][kx][group]
constexpr int NTHREADS = 192;                         // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue
constexpr int MAX_STAGES = 4;
constexpr int SMEM_LIMIT = 220 * 1024;

struct Args {
    float* dw;                                        // OIHW [Cout][Cin][3][3], accumulated with atomics
    int Cin, Cout;
    int groups;                                       // channel groups (of 8) per CTA: all of them, or 16 when blockIdx.y picks a block
    int mblocks;                                      // M blocks (128 channels) per CTA
    int nstages, sets;
    int H, W, tiles_x, tiles_y, n_tiles, tiles_per_cta;
    int g_grp0;                                       // first channel group of the gradient buffer (16-output-channel passes of the
                                                      // TransitionUp / first convolutions: pass p reads groups 2p, 2p + 1)
    int swap;                                         // 1: operands swapped (first convolution, Cin <= 8 and Cout <= 128): the M operand
                                                      // (`amap`) holds the GRADIENT planes (rows = co, no halo), the nine shifted boxes
                                                      // are taken from the ACTIVATION planes (`gmap`, one real group; the second group of
                                                      // a box is out of bounds = zeros): all output channels in ONE pass instead of
                                                      // Cout/16, D[co][tap * 16 + ci] with tap (ty, tx) = weight tap (2 - ty, 2 - tx)
};

__host__ __device__ inline int stage_bytes(int groups) { return groups * A_PLANE + G_BYTES; }
// the last M block reads 16 planes: past the activation planes into the gradient planes (finite numbers, rows never used),
// and past the end of the last stage when groups + 18 < 16 * mblocks -> pad
__host__ inline size_t smem_bytes(int groups, int mblocks, int nstages) {
    const long long over = 16ll * mblocks * A_PLANE - (long long)stage_bytes(groups);
    return 1024 + (size_t)nstages * stage_bytes(groups) + (over > 0 ? (size_t)over : 0);
}

// plane-major bf16 buffer [C/8 groups][B][H][W][8 channels] -> rank-4 map (8 W elements of an image row, H, B, groups) with a box
// of (8 * TW, box_h, 1, groups): one box line = 16 pixels x 8 channels = 256 contiguous bytes (the first version kept the
// by-products as [pixels][C] and fetched 16-byte lines: the TMA unit, not HBM, bounded the kernel at ~1 TB/s), and shared memory
// receives [group][row][pixel][8 channels] -- the MN-major SWIZZLE_NONE planes the descriptors address.  Columns left of 0 /
// right of W (x coordinate (x0 - 1) * 8 ...) and rows outside the image are zero-filled.
static inline bool make_map(CUtensorMap* map, const void* base, int B, int H, int W, int box_h, int groups_total, int groups) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)groups_total};
    const cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)B * H * W * 16};
    const cuuint32_t box[4] = {8u * TW, (cuuint32_t)box_h, 1u, (cuuint32_t)groups};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- operand packers for the layers whose data-gradient kernel does not leave the by-products behind ----------------------
// corrected output gradient G = g + A_c + B_c x (lazy BatchNorm term) of `Cout` channels of a level buffer -> plane-major bf16
// [Cout/8][npix][8] (padding channels zero), and its per-channel sum = the conv bias gradient.  4 lanes per pixel and 16-channel
// slice (blockIdx.y), HBM-bound.
__global__ void __launch_bounds__(256)
grad_pack16_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ ab, unsigned short* __restrict__ out,
                   float* __restrict__ db, int C, int off, int Cout, long long npix_per_group, int G) {
    pdl_enter();
    const int c0 = 16 * blockIdx.y;
    __shared__ float red[8][16];
    const int grp = threadIdx.x & 3, ch = c0 + grp * 4;
    const bool ch_ok = ch < Cout;
    const size_t npix = (size_t)npix_per_group * G;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int gi = 0; gi < G; ++gi) {
        float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0;
        if (ch_ok) {
            const float* abp = ab + ((size_t)gi * C + off + ch) * 2;
            k0 = __ldg(reinterpret_cast<const float4*>(abp)); k1 = __ldg(reinterpret_cast<const float4*>(abp + 4));
        }
        for (long long p = (long long)blockIdx.x * 64 + (threadIdx.x >> 2); p < npix_per_group; p += (long long)gridDim.x * 64) {
            const size_t pp = (size_t)gi * npix_per_group + p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch_ok) {
                const size_t o = pp * C + off + ch;
                const float4 gq = __ldg(reinterpret_cast<const float4*>(g + o)), xq = __ldg(reinterpret_cast<const float4*>(x + o));
                v.x = gq.x + fmaf(k0.y, xq.x, k0.x); v.y = gq.y + fmaf(k0.w, xq.y, k0.z);
                v.z = gq.z + fmaf(k1.y, xq.z, k1.x); v.w = gq.w + fmaf(k1.w, xq.w, k1.z);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            *reinterpret_cast<uint2*>(out + ((size_t)(ch >> 3) * npix + pp) * 8 + (ch & 7)) =
                make_uint2(tcwgrad_pack_bf16(v.x, v.y), tcwgrad_pack_bf16(v.z, v.w));
        }
    }
    if (!db) return;
#pragma unroll
    for (int o2 = 4; o2 < 32; o2 <<= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o2); s.y += __shfl_xor_sync(0xffffffffu, s.y, o2);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o2); s.w += __shfl_xor_sync(0xffffffffu, s.w, o2);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 4) { red[warp][lane * 4] = s.x; red[warp][lane * 4 + 1] = s.y; red[warp][lane * 4 + 2] = s.z; red[warp][lane * 4 + 3] = s.w; }
    __syncthreads();
    if (threadIdx.x < 16 && c0 + threadIdx.x < Cout) {
        float t = 0.f;
        for (int wq = 0; wq < 8; ++wq) t += red[wq][threadIdx.x];
        atomicAdd(db + c0 + threadIdx.x, t);
    }
}

// NCHW input images (<= 8 channels) -> plane-major bf16 [1][npix][8], channels past Cimg zero (operand of the first convolution's
// weight gradient, models.py:111-113)
__global__ void __launch_bounds__(256)
image_pack16_kernel(const float* __restrict__ img, unsigned short* __restrict__ out, int B, int Cimg, long long hw) {
    pdl_enter();
    const long long total = (long long)B * hw;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long b = i / hw, p = i - b * hw;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = c < Cimg ? __ldg(img + ((size_t)b * Cimg + c) * hw + p) : 0.f;
        *reinterpret_cast<uint4*>(out + (size_t)i * 8) = make_uint4(tcwgrad_pack_bf16(v[0], v[1]), tcwgrad_pack_bf16(v[2], v[3]),
                                                                    tcwgrad_pack_bf16(v[4], v[5]), tcwgrad_pack_bf16(v[6], v[7]));
    }
}

// nearest-neighbour x2 upsampling (models.py:73) of `cin` channels of a half-resolution level buffer -> plane-major bf16
// [cin/8][B*H*W][8] at the FULL resolution (operand of the TransitionUp convolution's weight gradient); thread = (pixel, group)
__global__ void __launch_bounds__(256)
upsample_pack16_kernel(const float* __restrict__ src, int srcC, int src_off, int cin, unsigned short* __restrict__ out, int B, int H, int W) {
    pdl_enter();
    const int ng = cin >> 3;
    const size_t npix = (size_t)B * H * W;
    const long long total = (long long)npix * ng;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int gq = (int)(i % ng);
        const size_t p = (size_t)(i / ng);
        const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((size_t)W * H));
        const float* sp = src + (((size_t)b * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * srcC + src_off + gq * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(sp)), c = __ldg(reinterpret_cast<const float4*>(sp + 4));
        *reinterpret_cast<uint4*>(out + ((size_t)gq * npix + p) * 8) = make_uint4(tcwgrad_pack_bf16(a.x, a.y), tcwgrad_pack_bf16(a.z, a.w),
                                                                                  tcwgrad_pack_bf16(c.x, c.y), tcwgrad_pack_bf16(c.z, c.w));
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
dense_wgrad_gemm_kernel(const Args A, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap gmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int stage = stage_bytes(A.groups);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                   // barriers first, stages from byte 1024
    unsigned char* stages = smem + 1024;
    uint64_t* full = bars;
    uint64_t* empty = bars + MAX_STAGES;
    uint64_t* accum = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    const int ntiles = t_end - t_begin;
    const int grp0 = blockIdx.y * A.groups;                               // first channel group of this CTA (0 unless blocks are split)

    pdl_trigger();
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accum, 1);
        tc::fence_mbar_init();
        tma::prefetch_map(&amap); tma::prefetch_map(&gmap);
    }
    pdl_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA issuer
        if (lane == 0) {
            const int per_img = A.tiles_x * A.tiles_y;
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % A.nstages;
                if (it >= A.nstages) tc::mbar_wait(empty + s, ((it / A.nstages) - 1) & 1);
                const int t = t_begin + it;
                const int b = t / per_img, rem = t - b * per_img;
                const int tx = rem / A.tiles_y, ty = rem - tx * A.tiles_y;     // column-major: consecutive tiles are vertical neighbours
                const int y0 = ty * TR, x0 = tx * TW;                           // (the gradient halo rows they share hit the L2)
                unsigned char* st = stages + (size_t)s * stage;
                tc::mbar_expect_tx(full + s, (uint32_t)stage);
                tma::load_4d(st, &amap, x0 * 8, y0, b, grp0, full + s);
                unsigned char* gs = st + (size_t)A.groups * A_PLANE;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
                    tma::load_4d(gs + tap * 2 * G_PLANE, &gmap, (x0 - (tap % 3 - 1)) * 8, y0 - (tap / 3 - 1), b, A.g_grp0, full + s);
                tc::mbar_arrive(full + s);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer: convergent, one elected lane issues
        const uint32_t idesc = tc::instr_desc(tc::FMT_BF16, 128, NB, 1, 1);               // both operands MN-major
        const uint64_t a_hi = tc::smem_desc(0, 128, A_PLANE), g_hi = tc::smem_desc(0, 128, G_PLANE);
        // everything the loop needs is kept incremental (the issuing thread's own instruction stream is the critical path)
        const uint32_t set_cols = (uint32_t)A.mblocks * NB;
        uint32_t set = 0, col = tmem, acc = 0;                     // accumulator set of the running K-step, its first column
        int s = 0, ph = 0;
        for (int it = 0; it < ntiles; ++it) {
            tc::mbar_wait(full + s, ph);
            tc::tc_fence_after();
            const uint32_t a_base = tc::smem_u32(stages + (size_t)s * stage), g_base = a_base + (uint32_t)A.groups * A_PLANE;
            uint64_t ad0 = a_hi | (uint64_t)(a_base >> 4), bd = g_hi | (uint64_t)(g_base >> 4);
#pragma unroll 1
            for (int k = 0; k < KPX / 16; ++k) {
                uint64_t ad = ad0;
                uint32_t cc = col;
                for (int mb = 0; mb < A.mblocks; ++mb) {
                    tc::mma_f16_w(cc, ad, bd, idesc, acc);
                    ad += (uint64_t)(16 * A_PLANE >> 4); cc += NB;
                }
                ad0 += 16; bd += 16;                                // 16 pixel rows of 16 bytes
                col += set_cols;
                if (++set == (uint32_t)A.sets) { set = 0; col = tmem; acc = 1; }
            }
            tc::tc_commit_w(empty + s);
            if (++s == A.nstages) { s = 0; ph ^= 1; }
        }
        tc::tc_commit_w(accum);
    } else {
        // ---------------------------------------------------------------- epilogue: D[mb][ky][kx*16 + co] -> atomicAdd into OIHW
        tc::mbar_wait(accum, 0);
        tc::tc_fence_after();
        const int q = warp & 3;
        const int total_k = ntiles * (KPX / 16);
        const int used_sets = total_k < A.sets ? total_k : A.sets;
        if (ntiles > 0) {
            for (int mb = 0; mb < A.mblocks; ++mb) {
                const int ci = grp0 * 8 + mb * 128 + q * 32 + lane;
#pragma unroll 1
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
                    for (int kx = 0; kx < 3; ++kx) {
                        float v[16];
                        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * NB + ky * 48 + kx * 16), v);
                        for (int set = 1; set < used_sets; ++set) {
                            float w[16];
                            tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * A.mblocks + mb) * NB + ky * 48 + kx * 16), w);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += w[j];
                        }
                        if (A.swap) {
                            if (ci < A.Cout) {                                   // the lane owns an OUTPUT channel, the 16 columns are ci
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (j < A.Cin) atomicAdd(A.dw + (((size_t)ci * A.Cin + j) * 3 + (2 - ky)) * 3 + (2 - kx), v[j]);
                            }
                        } else if (ci < A.Cin) {
#pragma unroll
                            for (int co = 0; co < 16; ++co)
                                if (co < A.Cout) atomicAdd(A.dw + (((size_t)co * A.Cin + ci) * 3 + ky) * 3 + kx, v[co]);
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc::tmem_dealloc(tmem, 512);
    }
}

}  // namespace tcwgrad3
}  // namespace endo
