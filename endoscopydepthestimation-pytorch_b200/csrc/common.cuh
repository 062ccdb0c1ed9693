// Shared device/host helpers for libendo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/endo_b200.h"

namespace endo {

extern unsigned long long g_launch_count;   // defined in api.cu

// Optional per-category device timing (endo_prof_* in api.cu): when enabled, every launch site brackets
// its kernel with two CUDA events on the launching stream.  Off by default (two predictable branches).
enum ProfCat { PC_CONV_DENSE_FWD = 0, PC_CONV_TRANS_FWD, PC_DGRAD, PC_WGRAD, PC_BN, PC_FINAL, PC_WARP, PC_FLOW,
               PC_SCALE, PC_LOSS, PC_OPT, PC_DGRAD_TRANS, PC_WGRAD_TRANS, PC_COUNT };   // PC_DGRAD / PC_WGRAD: DenseLayers only
extern int g_prof_on;
void prof_begin(int cat, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {
    cudaStream_t s; bool on;
    ProfScope(int cat, cudaStream_t st) : s(st), on(g_prof_on != 0) { if (on) prof_begin(cat, s); }
    ~ProfScope() { if (on) prof_end(s); }
};

#define ENDO_CHECK_LAUNCH()                                         \
    do {                                                            \
        ++endo::g_launch_count;                                     \
        if (cudaPeekAtLastError() != cudaSuccess) {                 \
            cudaGetLastError();                                     \
            return ENDO_ERR_CUDA;                                   \
        }                                                           \
    } while (0)

#define ENDO_CUDA(call)                                             \
    do {                                                            \
        if ((call) != cudaSuccess) {                                \
            cudaGetLastError();                                     \
            return ENDO_ERR_CUDA;                                   \
        }                                                           \
    } while (0)

// The opt-in for > 48 KB of dynamic shared memory is a PER-DEVICE function attribute: remember per device (bit mask over
// the current device ordinal; the header requires the current device to be the one that owns `stream`) which kernels
// have been configured.  Benign race: two threads may both set the same attribute to the same value.
#define ENDO_SET_MAX_SMEM(kern, bytes)                                                                     \
    do {                                                                                                   \
        static unsigned long long endo_cfg_mask_ = 0ull;                                                   \
        int endo_dev_ = 0;                                                                                 \
        ENDO_CUDA(cudaGetDevice(&endo_dev_));                                                              \
        if (endo_dev_ < 0 || endo_dev_ >= 64) return ENDO_ERR_NO_DEVICE;                                   \
        if (!((endo_cfg_mask_ >> endo_dev_) & 1ull)) {                                                     \
            ENDO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            endo_cfg_mask_ |= 1ull << endo_dev_;                                                           \
        }                                                                                                  \
    } while (0)

// Programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl() may become resident while its
// predecessor on the stream is still draining; it must run pdl_wait() BEFORE its first access to global memory (reads AND
// writes: the predecessor may still be reading what this kernel overwrites) and may do on-chip set-up (barrier init, TMEM
// allocation, descriptor prefetch) ahead of it.  pdl_trigger() at the top of a kernel lets ITS successor do the same as soon
// as every CTA of this kernel has started.  ENDO_PDL=0 launches everything with full stream serialisation (A/B switch).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ENDO_PDL"); v = e ? (atoi(e) != 0) : 1; }
    return v != 0;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200
constexpr float kBnEps = 1e-5f;       // nn.BatchNorm2d defaults (models.py:22,59)
constexpr float kBnMomentum = 0.1f;

// ---------------------------------------------------------------------------------------------
// Deterministic reductions.
//   Level 1: every block reduces N per-thread fp32 partials to N doubles (warp shuffles, fixed tree).
//   Level 2: blocks write their N doubles to a partial array; the block that arrives last (ticket
//            counter in the workspace header, self-resetting) sums the partials in a fixed order.
// The result therefore does not depend on block scheduling: bit-identical run to run.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int N, int THREADS>
__device__ __forceinline__ void block_sum(double (&v)[N], double* smem /* N * THREADS/32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) smem[i * NW + warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = lane < NW ? smem[i * NW + lane] : 0.0;
            s = warp_sum(s);
            v[i] = s;   // valid in every lane of warp 0
        }
    }
    __syncthreads();
}

// returns true in every thread of the block that arrived last among `total` blocks on `counter`
__device__ __forceinline__ bool arrive_is_last(unsigned* counter, unsigned total) {
    __shared__ unsigned s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(counter, 1u);
        s_last = (t == total - 1u) ? 1u : 0u;
        if (s_last) *counter = 0u;   // self-reset: header is zero again when the kernel ends
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0u;
}

__device__ __forceinline__ double ld_cg(const double* p) {
    return __ldcg(p);   // bypass L1: partials were written by other SMs
}

}  // namespace endo
