// Library-level entry points of libendo_b200.so (version, error strings, launch counter).
#include "common.cuh"

namespace endo {
unsigned long long g_launch_count = 0ull;
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
