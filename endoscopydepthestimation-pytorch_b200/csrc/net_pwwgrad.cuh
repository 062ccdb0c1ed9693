// Weight gradient of the TransitionDown 1x1 convolution (reference models.py:56-67) as ONE real C x C GEMM on tcgen05:
//
//   dW[co][ci] = sum_p R[p][co] * act[p][ci]        M = ci (TMEM lanes, blocks of 128), N = co, K = pixels
//
// Both operands are bf16 [pixels][C] matrices that the data-gradient kernel of the same layer (tcpw::pw_gemm_kernel, mode 1)
// leaves behind as by-products: R = the pooled gradient routed to the argmax position with the lazy BatchNorm term (its
// operand A), act = relu(bn(x)) (its epilogue evaluates exactly that for the ReLU mask).  So this kernel has NO transform
// warps: one thread streams (8 channels x KT pixels x C/8 groups) boxes with cp.async.bulk.tensor -- the rank-3 tensor map
// (channel-in-group, pixel, group) lands a box directly as the MN-major SWIZZLE_NONE planes the MMA descriptors address --
// one thread issues the MMAs, the accumulators (<= 512 TMEM columns) stay resident over all pixel tiles of the CTA, and four
// warps add them into dW at the end.  Round 1/early round 2 ran the DenseLayer weight-gradient kernel in a 1x1 mode instead:
// 48 output channels per pass and 64-channel input blocks, i.e. (C/48) x ceil(C/64) passes over activations and gradients
// (4 for C = 96: 1.0 ms at the first TransitionDown of a 16 x 256 x 320 batch; this kernel reads 2 x 252 MB once).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace endo {
namespace tcpww {

constexpr int NTHREADS = 192;                  // warp 0: TMA, warp 1: MMA, warps 2-5: epilogue (TMEM lane quadrant = warp & 3)
constexpr int MAX_STAGES = 4;
constexpr int SMEM_LIMIT = 220 * 1024;

struct Args {
    float* dw;                                 // [C][C] (OIHW with 1x1 taps), accumulated with atomics
    int C;                                     // channels (multiple of 16)
    int Nper;                                  // output channels per blockIdx.y (multiple of 16, <= 256)
    int KT;                                    // pixels per stage: 128 or 64
    int nstages;
    int n_tiles, tiles_per_cta;
    int sets;                                  // accumulator sets rotated over the K-steps (an MMA that accumulates into the columns
                                               // its predecessor wrote waits for it: independent chains keep the pipe busy)
};

__host__ __device__ inline int stage_bytes(int C, int Nper, int KT) { return (C + Nper) * KT * 2; }
// An M = 128 MMA reads 16 channel-group planes from its start plane; when the last block of C has fewer, the read runs on
// into the R planes of the stage (finite numbers; those accumulator rows are never used) -- and past the end of the LAST stage
// when C/8 + Nper/8 < 16 * blocks: pad the allocation so that it stays inside it.
__host__ inline size_t smem_bytes(int C, int Nper, int KT, int nstages) {
    const int mblocks = (C + 127) / 128;
    const int over = 16 * mblocks - (C / 8 + Nper / 8);
    return 1024 + (size_t)nstages * stage_bytes(C, Nper, KT) + (over > 0 ? (size_t)over * KT * 16 : 0);
}

// bf16 matrix [P][C] -> rank-3 map (8 channels, P pixels, C/8 groups) with a box of (8, KT, groups): shared memory receives
// [group][pixel][8 channels] = one 16-byte row per pixel and channel group, planes KT * 16 bytes apart.
static inline bool make_planes_map(CUtensorMap* map, const void* base, long long P, int C, int KT, int groups) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[3] = {8u, (cuuint64_t)P, (cuuint64_t)(C / 8)};
    const cuuint64_t gstr[2] = {(cuuint64_t)C * 2, 16u};
    const cuuint32_t box[3] = {8u, (cuuint32_t)KT, (cuuint32_t)groups};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            tc::smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
pw_wgrad_kernel(const Args A, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap rmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int stage = stage_bytes(A.C, A.Nper, A.KT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                   // barriers first, stages from byte 1024
    unsigned char* stages = smem + 1024;
    uint64_t* full = bars;                     // [MAX_STAGES] 1 arrival + transaction bytes
    uint64_t* empty = bars + MAX_STAGES;       // [MAX_STAGES] tcgen05.commit
    uint64_t* accum = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t_begin = blockIdx.x * A.tiles_per_cta;
    const int t_end = min(t_begin + A.tiles_per_cta, A.n_tiles);
    const int ntiles = t_end - t_begin;
    const int n0 = blockIdx.y * A.Nper;
    const int mblocks = (A.C + 127) >> 7;

    pdl_trigger();
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        for (int i = 0; i < MAX_STAGES; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accum, 1);
        tc::fence_mbar_init();
        tma::prefetch_map(&amap); tma::prefetch_map(&rmap);
    }
    pdl_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA issuer
        if (lane == 0) {
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % A.nstages;
                if (it >= A.nstages) tc::mbar_wait(empty + s, ((it / A.nstages) - 1) & 1);
                unsigned char* st = stages + (size_t)s * stage;
                tc::mbar_expect_tx(full + s, (uint32_t)stage);
                load_3d(st, &amap, 0, (t_begin + it) * A.KT, 0, full + s);                                 // act: all C/8 groups
                load_3d(st + (size_t)A.C * A.KT * 2, &rmap, 0, (t_begin + it) * A.KT, n0 >> 3, full + s);    // R: Nper/8 groups from n0
                tc::mbar_arrive(full + s);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer: convergent, one elected lane issues
        const uint32_t idesc = tc::instr_desc(tc::FMT_BF16, 128, A.Nper, 1, 1);          // both operands MN-major
        const uint32_t plane = (uint32_t)A.KT * 16u;
        const uint64_t d_hi = tc::smem_desc(0, 128, plane);                               // LBO = 8-pixel groups, SBO = channel-group planes
        const int ksteps = A.KT >> 4;
        // incremental bookkeeping: the issuing thread's own instruction stream is the critical path of this kernel
        const uint32_t set_cols = (uint32_t)(mblocks * A.Nper);
        uint32_t set = 0, col = tmem, acc = 0;
        int s = 0, ph = 0;
        for (int it = 0; it < ntiles; ++it) {
            tc::mbar_wait(full + s, ph);
            tc::tc_fence_after();
            const uint32_t a_base = tc::smem_u32(stages + (size_t)s * stage), b_base = a_base + (uint32_t)A.C * A.KT * 2u;
            uint64_t ad0 = d_hi | (uint64_t)(a_base >> 4), bd = d_hi | (uint64_t)(b_base >> 4);
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k) {
                uint64_t ad = ad0;
                uint32_t cc = col;
                for (int mb = 0; mb < mblocks; ++mb) {
                    tc::mma_f16_w(cc, ad, bd, idesc, acc);
                    ad += (uint64_t)(plane); cc += (uint32_t)A.Nper;           // 16 planes = 16 * plane bytes = plane 16-byte units
                }
                ad0 += 16; bd += 16;
                col += set_cols;
                if (++set == (uint32_t)A.sets) { set = 0; col = tmem; acc = 1; }
            }
            tc::tc_commit_w(empty + s);
            if (++s == A.nstages) { s = 0; ph ^= 1; }
        }
        tc::tc_commit_w(accum);
    } else {
        // ---------------------------------------------------------------- epilogue: sum the accumulator sets, add into dW
        tc::mbar_wait(accum, 0);
        tc::tc_fence_after();
        const int q = warp & 3;
        const int total_k = ntiles * (A.KT >> 4);
        const int used_sets = total_k < A.sets ? total_k : A.sets;                        // sets that were ever written
        if (ntiles > 0) {
            for (int mb = 0; mb < mblocks; ++mb) {
                const int ci = mb * 128 + q * 32 + lane;
                for (int c16 = 0; c16 < A.Nper; c16 += 16) {
                    float v[16];
                    tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * A.Nper + c16), v);
                    for (int set = 1; set < used_sets; ++set) {
                        float w[16];
                        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * mblocks + mb) * A.Nper + c16), w);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += w[j];
                    }
                    if (ci < A.C) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int co = n0 + c16 + j;
                            if (co < A.C) atomicAdd(A.dw + (size_t)co * A.C + ci, v[j]);
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

}  // namespace tcpww
}  // namespace endo
