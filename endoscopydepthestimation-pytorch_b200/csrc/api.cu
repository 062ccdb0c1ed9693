// Library-level entry points of libendo_b200.so (version, error strings, launch counter).
#include "common.cuh"

#include <vector>

namespace endo {
unsigned long long g_launch_count = 0ull;
int g_prof_on = 0;

struct ProfRec { int cat; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static int g_prof_open = -1;

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_begin(int cat, cudaStream_t s) {
    ProfRec r; r.cat = cat; r.a = prof_event(); r.b = prof_event();
    cudaEventRecord(r.a, s);
    g_prof_recs.push_back(r);
    g_prof_open = (int)g_prof_recs.size() - 1;
}
void prof_end(cudaStream_t s) {
    if (g_prof_open >= 0) cudaEventRecord(g_prof_recs[g_prof_open].b, s);
    g_prof_open = -1;
}
}  // namespace endo

extern "C" void endo_prof_enable(int on) { endo::g_prof_on = on; }
extern "C" int endo_prof_categories(void) { return endo::PC_COUNT; }
extern "C" const char* endo_prof_category_name(int c) {
    static const char* names[] = {"conv_dense_fwd", "conv_trans_fwd", "conv_dense_dgrad", "conv_dense_wgrad", "bn_bookkeeping",
                                  "final_conv", "depth_warp", "flow_from_depth", "depth_scale", "losses", "optimizer",
                                  "conv_trans_dgrad", "conv_trans_wgrad"};
    return (c >= 0 && c < endo::PC_COUNT) ? names[c] : "?";
}
// Synchronises the device, adds every recorded launch's duration to ms[cat] / counts[cat] and clears the records.
extern "C" int endo_prof_collect(double* ms, unsigned long long* counts) {
    if (cudaDeviceSynchronize() != cudaSuccess) return ENDO_ERR_CUDA;
    for (auto& r : endo::g_prof_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cat >= 0 && r.cat < endo::PC_COUNT) {
            if (ms) ms[r.cat] += (double)t;
            if (counts) counts[r.cat] += 1ull;
        }
        endo::g_prof_pool.push_back(r.a); endo::g_prof_pool.push_back(r.b);
    }
    endo::g_prof_recs.clear();
    return ENDO_OK;
}

extern "C" int endo_version(void) { return 101; }

extern "C" unsigned long long endo_launch_count(void) { return endo::g_launch_count; }

extern "C" const char* endo_strerror(int code) {
    switch (code) {
        case ENDO_OK: return "ok";
        case ENDO_ERR_BAD_SHAPE: return "bad shape (dims must be positive; the network needs H, W multiples of 2^n_down)";
        case ENDO_ERR_BAD_POINTER: return "null or misaligned pointer";
        case ENDO_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
        case ENDO_ERR_CUDA: return "CUDA runtime / launch error";
        case ENDO_ERR_CONFIG: return "unsupported network configuration";
        case ENDO_ERR_NO_DEVICE: return "no sm_100 device";
        default: return "unknown error";
    }
}
