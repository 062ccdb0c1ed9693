// tcgen05 bring-up probe: D[128][N] = A[shift .. shift+128][K] * B[N][K]^T on the 5th-generation tensor
// cores, with the operands staged by ordinary threads into the SWIZZLE_NONE canonical layouts that the
// convolution kernels use (K-major with 16-byte row pitch, or MN-major), tf32 or bf16, fp32 accumulation
// in TMEM.  It exists so that every descriptor convention the conv kernels rely on (row sliding through
// the start address, LBO/SBO meaning, instruction-descriptor fields, TMEM lane/column mapping) is checked
// against a plain matmul on real hardware by tests/test_gpu_tc_probe.py.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace endo {

struct ProbeArgs {
    const float* A; const float* B; float* D;
    int a_rows, N, K, shift, fmt, a_mn, b_mn;
    int swz, reps; long long* cycles;      // swz: 0 = SWIZZLE_NONE, 2 = SWIZZLE_128B (K-major only); reps: repeat the MMA chain (timing)
    int rotate;                            // timing only: consecutive reps target `rotate` different accumulator tiles
};

__global__ void __launch_bounds__(128)
tc_probe_kernel(const ProbeArgs P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int es = (P.fmt == tc::FMT_TF32) ? 4 : 2;       // element size in shared memory
    const int T = 16 / es;                                // elements per 16-byte chunk
    const int kstep = 32 / es;                            // K per MMA: 8 (tf32) or 16 (bf16)
    // ---- shared layout
    unsigned char* a_s = smem;
    const uint32_t a_bytes = (uint32_t)P.a_rows * P.K * es;
    unsigned char* b_s = smem + ((a_bytes + 1023) / 1024) * 1024;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    if (!P.a_mn) { a_sbo = 128; a_lbo = (uint32_t)P.a_rows * 16; }         // K-major: planes of rows x 16 B
    else { a_sbo = 128; a_lbo = (uint32_t)(P.a_rows / T) * 128; }          // MN-major: 128-B blocks (T mn x 8 k)
    if (!P.b_mn) { b_sbo = 128; b_lbo = (uint32_t)P.N * 16; }
    else { b_sbo = 128; b_lbo = (uint32_t)(P.N / T) * 128; }

    const int kpr = 128 / es;                             // K elements per 128-byte swizzled row
    auto put = [&](unsigned char* base, bool mn, uint32_t lbo, uint32_t sbo, int r, int k, float v) {
        uint32_t off;
        if (P.swz == 2) {                                 // slab (k / kpr) of [rows][128 B], chunk index XOR (row & 7)
            const int rows = (base == a_s) ? P.a_rows : P.N;
            const int kk = k % kpr;
            off = (uint32_t)(k / kpr) * (uint32_t)rows * 128u + (uint32_t)r * 128u +
                  (uint32_t)((((kk * es) >> 4) ^ (r & 7)) << 4) + (uint32_t)((kk * es) & 15);
        } else if (!mn) off = (uint32_t)(k / T) * lbo + (uint32_t)r * 16 + (uint32_t)(k % T) * es;
        else off = (uint32_t)(r % T) * es + (uint32_t)(k % 8) * 16 + (uint32_t)(r / T) * sbo + (uint32_t)(k / 8) * lbo;
        if (es == 4) *reinterpret_cast<float*>(base + off) = v;
        else *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
    };
    for (int i = tid; i < P.a_rows * P.K; i += 128) put(a_s, P.a_mn, a_lbo, a_sbo, i / P.K, i % P.K, P.A[i]);
    for (int i = tid; i < P.N * P.K; i += 128) put(b_s, P.b_mn, b_lbo, b_sbo, i / P.K, i % P.K, P.B[i]);

    const uint32_t ncols = P.rotate > 1 ? 512u : (P.N <= 32 ? 32 : (P.N <= 64 ? 64 : (P.N <= 128 ? 128 : 256)));
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = tc::instr_desc(P.fmt, 128, P.N, P.a_mn, P.b_mn);
        const long long t0 = clock64();
        for (int rep = 0; rep < P.reps; ++rep) {
            for (int k0 = 0; k0 < P.K; k0 += kstep) {
                uint32_t a_addr = tc::smem_u32(a_s), b_addr = tc::smem_u32(b_s);
                uint64_t ad, bd;
                if (P.swz == 2) {
                    a_addr += (uint32_t)(k0 / kpr) * (uint32_t)P.a_rows * 128u + (uint32_t)P.shift * 128u + (uint32_t)((k0 % kpr) * es);
                    b_addr += (uint32_t)(k0 / kpr) * (uint32_t)P.N * 128u + (uint32_t)((k0 % kpr) * es);
                    ad = tc::smem_desc_sw128(a_addr); bd = tc::smem_desc_sw128(b_addr);
                } else {
                    if (!P.a_mn) a_addr += (uint32_t)(k0 / T) * a_lbo + (uint32_t)P.shift * 16;
                    else a_addr += (uint32_t)(k0 / 8) * a_lbo + (uint32_t)(P.shift / T) * a_sbo;
                    if (!P.b_mn) b_addr += (uint32_t)(k0 / T) * b_lbo;
                    else b_addr += (uint32_t)(k0 / 8) * b_lbo;
                    ad = tc::smem_desc(a_addr, a_lbo, a_sbo); bd = tc::smem_desc(b_addr, b_lbo, b_sbo);
                }
                const uint32_t dcol = (P.rotate > 1) ? (uint32_t)((rep % P.rotate) * P.N) : 0u;
                const uint32_t accf = (P.rotate > 1) ? (uint32_t)(rep >= P.rotate || k0 > 0) : (uint32_t)((rep | k0) != 0);
                if (P.fmt == tc::FMT_TF32) tc::mma_tf32(tmem + dcol, ad, bd, idesc, accf);
                else tc::mma_f16(tmem + dcol, ad, bd, idesc, accf);
            }
        }
        tc::tc_commit(&bar);
        tc::mbar_wait(&bar, 0);
        if (P.cycles) P.cycles[0] = clock64() - t0;
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < P.N; c0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 8; ++j) P.D[(size_t)(warp * 32 + lane) * P.N + c0 + j] = v[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace endo

using namespace endo;

extern "C" int endo_tc_probe(const float* A, const float* B, float* D, int a_rows, int N, int K, int shift, int fmt,
                             int a_mn_major, int b_mn_major, int swizzle, int reps, long long* cycles, int rotate,
                             endo_stream_t stream) {
    if (!A || !B || !D) return ENDO_ERR_BAD_POINTER;
    const int kstep = (fmt == tc::FMT_TF32) ? 8 : 16;
    const int T = (fmt == tc::FMT_TF32) ? 4 : 8;
    if ((fmt != tc::FMT_TF32 && fmt != tc::FMT_BF16) || N < 16 || N > 256 || (N % 16) || K < kstep || (K % kstep) ||
        shift < 0 || a_rows < 128 + shift || (a_rows % 8) || (a_mn_major && ((shift % T) || (a_rows % T))))
        return ENDO_ERR_BAD_SHAPE;
    const int es = (fmt == tc::FMT_TF32) ? 4 : 2;
    if (swizzle != 0 && swizzle != 2) return ENDO_ERR_BAD_SHAPE;
    if (swizzle == 2 && (a_mn_major || b_mn_major || (K * es) % 128)) return ENDO_ERR_BAD_SHAPE;
    if (reps < 1) reps = 1;
    if (rotate < 1) rotate = 1;
    if (rotate > 1 && rotate * N > 512) return ENDO_ERR_BAD_SHAPE;
    const size_t smem = ((size_t)a_rows * K * es + 1023) / 1024 * 1024 + (size_t)N * K * es + 2048;
    if (smem > 200 * 1024) return ENDO_ERR_BAD_SHAPE;
    ENDO_SET_MAX_SMEM(tc_probe_kernel, 200 * 1024);
    ProbeArgs p{A, B, D, a_rows, N, K, shift, fmt, a_mn_major, b_mn_major, swizzle, reps, cycles, rotate};
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

// ------------------------------------------------------------------------------------------------ TMA probe
// out[box_h][box_w][box_c] = the box of an NHWC buffer at (c0, x0, y0, b), loaded by ONE cp.async.bulk.tensor (zero fill
// outside the tensor): checks the tensor-map conventions the convolution kernels rely on (tests/test_gpu_tc_probe.py).
#include "tma.cuh"
namespace endo {
__global__ void __launch_bounds__(128)
tma_probe_kernel(const __grid_constant__ CUtensorMap map, float* __restrict__ out, int c0, int x0, int y0, int b, int n) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(&bar, (uint32_t)n * 4u);
        tma::load_4d(smem, &map, c0, x0, y0, b, &bar);
        tc::mbar_arrive(&bar);
    }
    tc::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<const float*>(smem)[i];
}
}  // namespace endo

extern "C" int endo_tma_probe(const float* src, int B, int H, int W, int C, int box_c, int box_w, int box_h, int c0, int x0,
                              int y0, int b, float* out, endo_stream_t stream) {
    if (!src || !out) return ENDO_ERR_BAD_POINTER;
    if (C % 4 || box_c % 4 || box_c > 256 || box_w > 256 || box_h > 256) return ENDO_ERR_BAD_SHAPE;
    CUtensorMap map;
    if (!endo::tma::make_nhwc_map(&map, src, B, H, W, C, box_c, box_w, box_h)) return ENDO_ERR_CUDA;
    const int n = box_c * box_w * box_h;
    if ((size_t)n * 4 > 200 * 1024) return ENDO_ERR_BAD_SHAPE;
    ENDO_SET_MAX_SMEM(endo::tma_probe_kernel, 200 * 1024);
    endo::tma_probe_kernel<<<1, 128, (size_t)n * 4, (cudaStream_t)stream>>>(map, out, c0, x0, y0, b, n);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}
