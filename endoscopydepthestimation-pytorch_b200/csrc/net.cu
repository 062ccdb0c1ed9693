// FCDenseNet forward / backward entry points (reference models.py:100-187 and the autograd backward
// PyTorch derives from it), fp32 FFMA path.  Host side: walks the NetPlan and enqueues the kernels of
// net_kernels.cuh on the caller's stream; no allocation, no synchronisation.
#include <cstdlib>
#include <mutex>
#include <vector>
#include "net_kernels.cuh"
#include "net_plan.cuh"
#include "net_tc.cuh"
#include "net_pw.cuh"
#include "net_wgrad2.cuh"
#include "net_fwd2.cuh"
#include "net_pwwgrad.cuh"
#include "net_wgrad3.cuh"

namespace endo {

// ------------------------------------------------------------------------------------------------ launchers
template <int KS, int PX, int CO, int NW, int LM, int EM, int WM, bool UP, int KCT = KC, int MINB = 1>
static int launch_conv(const ConvArgs& a, cudaStream_t s) {
    constexpr size_t smem = conv_smem_bytes<KS, PX, CO, NW, KCT>();
    auto kern = conv_kernel<KS, PX, CO, NW, LM, EM, WM, UP, KCT, MINB>;
    ENDO_SET_MAX_SMEM(kern, (int)smem);
    const int tiles = cdiv(a.ow, 32) * cdiv(a.oh, NW * PX);
    dim3 grid(tiles, cdiv(a.N, CO), a.B);
    ProfScope prof(WM == WM_DGRAD ? ((LM == LM_GRADPOOL || EM == EM_DGRAD_UP) ? PC_DGRAD_TRANS : PC_DGRAD)
                                  : ((LM == LM_BNRELU && EM == EM_STORE) ? PC_CONV_DENSE_FWD : PC_CONV_TRANS_FWD), s);
    launch_pdl(kern, grid, NW * 32, smem, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

// DenseLayer forward, software-pipelined variant (4-channel chunks, register prefetch, coefficient table in shared memory)
template <int PX, int CO>
static int launch_conv_pf(const ConvArgs& a, cudaStream_t s) {
    constexpr size_t base = conv_smem_bytes<3, PX, CO, 4, 4>();
    constexpr int kMaxK = 1536;
    auto kern = conv_kernel<3, PX, CO, 4, LM_BNRELU, EM_STORE, WM_FWD, false, 4, 3, true>;
    ENDO_SET_MAX_SMEM(kern, (int)(base + 16 * kMaxK));
    if (a.K > kMaxK) return ENDO_ERR_BAD_SHAPE;
    dim3 grid(cdiv(a.ow, 32) * cdiv(a.oh, 4 * PX), cdiv(a.N, CO), a.B);
    ProfScope prof(PC_CONV_DENSE_FWD, s);
    launch_pdl(kern, grid, 128, base + 16 * (size_t)a.K, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

template <int KS, int PX, int CO, int NW, int LM>
static int launch_conv_splitk(const ConvArgs& a, cudaStream_t s) {
    constexpr size_t smem = conv_smem_bytes<KS, PX, CO, NW>();
    auto kern = conv_kernel<KS, PX, CO, NW, LM, EM_PARTIAL, WM_FWD, false>;
    ENDO_SET_MAX_SMEM(kern, (int)smem);
    dim3 grid(cdiv(a.ow, 32) * cdiv(a.oh, NW * PX), a.ksplit, a.B);
    ProfScope prof(PC_CONV_DENSE_FWD, s);
    launch_pdl(kern, grid, NW * 32, smem, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

template <int KS, int CW, int NCG, int NPS, int LMA, int LMG, bool UP>
static int launch_wgrad(WgradArgs a, cudaStream_t s) {
    constexpr size_t smem = wgrad_smem_bytes<KS, CW, NCG>();
    auto kern = wgrad_kernel<KS, CW, NCG, NPS, LMA, LMG, UP>;
    ENDO_SET_MAX_SMEM(kern, (int)smem);
    const int ychunks = cdiv(a.a_K, 32), zchunks = cdiv(a.g_K, CW * NCG);
    a.n_tiles = a.B * cdiv(a.oh, 8) * cdiv(a.ow, 32);
    int want = (3 * kNumSMs) / (ychunks * zchunks);
    if (want < 1) want = 1;
    if (want > a.n_tiles) want = a.n_tiles;
    a.tiles_per_cta = cdiv(a.n_tiles, want);
    dim3 grid(cdiv(a.n_tiles, a.tiles_per_cta), ychunks, zchunks);
    ProfScope prof((KS == 3 && LMA == LM_BNRELU) ? PC_WGRAD : PC_WGRAD_TRANS, s);
    launch_pdl(kern, grid, KS * NCG * NPS * 32, smem, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

template <int KS, int CW, int NCG, int NPS, int LMA, int LMG, bool UP>
static int launch_wgrad2(WgradArgs a, cudaStream_t s) {
    if (a.G > 2) return launch_wgrad<KS, CW, NCG, NPS, LMA, LMG, UP>(a, s);     // coefficient cache holds two statistic groups
    constexpr size_t smem = wgrad2_smem_bytes<KS, CW, NCG>();
    auto kern = wgrad2_kernel<KS, CW, NCG, NPS, LMA, LMG, UP>;
    ENDO_SET_MAX_SMEM(kern, (int)smem);
    const int ychunks = cdiv(a.a_K, 32), zchunks = cdiv(a.g_K, CW * NCG);
    a.n_tiles = a.B * cdiv(a.oh, 8) * cdiv(a.ow, 32);
    int want = (2 * kNumSMs) / (ychunks * zchunks);
    if (want < 1) want = 1;
    if (want > a.n_tiles) want = a.n_tiles;
    a.tiles_per_cta = cdiv(a.n_tiles, want);
    dim3 grid(cdiv(a.n_tiles, a.tiles_per_cta), ychunks, zchunks);
    ProfScope prof((KS == 3 && LMA == LM_BNRELU) ? PC_WGRAD : PC_WGRAD_TRANS, s);
    launch_pdl(kern, grid, KS * NCG * NPS * 32, smem, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

// ENDO_TC_DISABLE (bit mask, debugging / A-B tests only): 1 = forward, 2 = data gradient, 4 = weight gradient fall back
// to the FFMA kernels even when the math mode asks for tensor cores (8/16/32/64: transition layers); 128 = no split-K
// in the FFMA DenseLayer forward; 8192 = no side stream; 16384 / 32768 = round-1 weight-gradient / forward kernels;
// 65536 = TransitionDown max-pool as a separate pass; 131072 = TransitionDown weight gradient through the 1x1 mode of the
// DenseLayer kernel; 262144 = DenseLayer / TransitionUp / first-convolution weight gradients through the round-2a kernels
// (no bf16 by-product planes); 524288 = first convolution forward on the FFMA kernel.  ENDO_PDL=0: no programmatic dependent launch.
static int tc_debug_mask() {
    const char* e = getenv("ENDO_TC_DEBUG");
    return e ? atoi(e) : 0;
}
static int tc_disable_mask() {
    const char* e = getenv("ENDO_TC_DISABLE");
    return e ? atoi(e) : 0;
}

static inline bool is_tc(int math) {
    return math == ENDO_MATH_TF32 || math == ENDO_MATH_TF32X3 || math == ENDO_MATH_BF16X3 || math == ENDO_MATH_BF16;
}
// forward operand scheme of the tensor-core modes: 0 = plain tf32, 1 = 3xTF32, 2 = bf16x3, 3 = plain bf16 (the bf16x3 staging,
// only the first-term product is issued: activations and weights rounded to bf16, fp32 accumulation -- BASELINE config 3)
static inline int x3_mode(int math) {
    return math == ENDO_MATH_TF32X3 ? 1 : (math == ENDO_MATH_BF16X3 ? 2 : (math == ENDO_MATH_BF16 ? 3 : 0));
}

// Weight-gradient kernels only feed the optimiser: they are enqueued on a side stream (forked from / joined to the caller's
// stream with events) so that the CTAs of one kernel fill the SMs the other leaves idle in its last wave and prologue (every
// tensor-core kernel here occupies a whole SM per CTA; 5-30 % of a launch is tail).  The GEMM weight-gradient kernels consume
// bf16 operand planes that the data-gradient kernel of the same layer writes: two plane sets alternate, Ctx::acquire_buf /
// release_buf order "GEMM of layer i has read set k" before "data gradient of layer i-2 rewrites set k" with events.  ENDO_NET_SINGLE_STREAM (math flag) or ENDO_TC_DISABLE bit 8192
// keeps everything on the caller's stream.
//
// Re-entrancy: a call BORROWS a (stream, fork event, join event) triple from a per-device pool for the duration of its
// host-side enqueue and returns it on every exit path (SideLease), after recording the join; concurrent callers (other host
// threads / streams) get different triples, so nobody re-records an event another caller is about to wait on.  Re-use by a
// LATER call is safe: cudaStreamWaitEvent captures the event's most recent record at the time of the call.  The triples are
// created on first use (not during stream capture: run one warm-up step before capturing a CUDA graph) and live until exit.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr, buf[2] = {nullptr, nullptr}; };
static std::mutex g_side_mu;
static std::vector<SideStream*> g_side_free[64];
static SideStream* side_acquire(int dev) {
    {
        std::lock_guard<std::mutex> lk(g_side_mu);
        auto& v = g_side_free[dev];
        if (!v.empty()) { SideStream* t = v.back(); v.pop_back(); return t; }
    }
    SideStream* t = new SideStream();
    if (cudaStreamCreateWithFlags(&t->s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->buf[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->buf[1], cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (t->buf[0]) cudaEventDestroy(t->buf[0]);
        if (t->join) cudaEventDestroy(t->join);
        if (t->fork) cudaEventDestroy(t->fork);
        if (t->s) cudaStreamDestroy(t->s);
        delete t;
        return nullptr;
    }
    return t;
}
// joins the side stream into the caller's stream and returns the triple to the pool when it goes out of scope
struct SideLease {
    SideStream* t = nullptr; int dev = 0; cudaStream_t caller = nullptr; bool forked = false;
    ~SideLease() {
        if (!t) return;
        if (forked) {           // every exit path: the caller's stream waits for whatever was enqueued on the side stream
            if (cudaEventRecord(t->join, t->s) != cudaSuccess || cudaStreamWaitEvent(caller, t->join, 0) != cudaSuccess)
                cudaGetLastError();
        }
        std::lock_guard<std::mutex> lk(g_side_mu);
        g_side_free[dev].push_back(t);
    }
};

struct Ctx {
    const NetPlan& P;
    char* acts; char* scratch;
    const float* params; float* gparams; float* bnbuf;
    cudaStream_t s;
    int training;
    int math;
    cudaStream_t sw = nullptr;        // stream of the weight-gradient kernels (== s when the side stream is off)
    cudaEvent_t ev_fork = nullptr;
    SideLease* lease = nullptr;
    // by-product buffer sets of the DenseLayer weight-gradient GEMM: layer i's data gradient (main stream) writes set i & 1, its
    // GEMM (side stream) reads it; ev_buf[k] = "the GEMM that last read set k has finished", awaited before set k is rewritten
    cudaEvent_t ev_buf[2] = {nullptr, nullptr};
    mutable int wg_layers = 0;
    mutable bool buf_busy[2] = {false, false};
    // next by-product set; the MAIN stream waits until the GEMM that last read it has finished
    int acquire_buf(int* bk) const {
        *bk = wg_layers++ & 1;
        if (sw != s && buf_busy[*bk]) ENDO_CUDA(cudaStreamWaitEvent(s, ev_buf[*bk], 0));
        return ENDO_OK;
    }
    int release_buf(int bk) const {                   // after the last GEMM that reads set bk was enqueued on sw
        if (sw != s) { ENDO_CUDA(cudaEventRecord(ev_buf[bk], sw)); buf_busy[bk] = true; }
        return ENDO_OK;
    }
    unsigned short* A16(int bk) const { return reinterpret_cast<unsigned short*>(scratch + P.a16_off[bk]); }
    unsigned short* G16(int bk) const { return reinterpret_cast<unsigned short*>(scratch + P.g16_off[bk]); }
    // everything enqueued on s so far happens-before what is enqueued on sw from now on
    int fork() const {
        if (sw == s) return ENDO_OK;
        ENDO_CUDA(cudaEventRecord(ev_fork, s));
        ENDO_CUDA(cudaStreamWaitEvent(sw, ev_fork, 0));
        lease->forked = true;
        return ENDO_OK;
    }
    float* X(int l) const { return reinterpret_cast<float*>(acts + P.x_off[l]); }
    double* ST(int l) const { return reinterpret_cast<double*>(acts + P.stat_off[l]); }
    float* MI(int l) const { return reinterpret_cast<float*>(acts + P.mi_off[l]); }
    float* COEF(const BnP& b) const { return reinterpret_cast<float*>(acts) + b.coef; }
    float* WPACK() const { return reinterpret_cast<float*>(acts + P.wpack_off); }
    float* WPACK_BWD() const { return reinterpret_cast<float*>(scratch + P.wpack_bwd_off); }
    float* GX(int l) const { return reinterpret_cast<float*>(scratch + P.gx_off[l]); }
    float* AB(int l) const { return reinterpret_cast<float*>(scratch + P.ab_off[l]); }
    double* BNRED() const { return reinterpret_cast<double*>(scratch + P.bnred_off); }
    double count(int l) const { return (double)(P.B / P.G) * P.h[l] * P.w[l]; }
};

static int bn_prepare(const Ctx& c, const BnP& bn, int level, int ch_off) {
    BnPrepArgs a;
    a.stats = c.ST(level); a.mi = c.MI(level); a.coef = c.COEF(bn);
    a.gamma = c.params + bn.gamma; a.beta = c.params + bn.beta;
    a.rmean = c.bnbuf + bn.rmean; a.rvar = c.bnbuf + bn.rvar;
    a.C = bn.c; a.Ctot = c.P.Ctot[level]; a.ch_off = ch_off; a.G = c.P.G; a.training = c.training;
    a.count = c.count(level);
    ProfScope prof(PC_BN, c.s);
    launch_pdl(bn_prepare_kernel, cdiv(bn.c, 128), 128, 0, c.s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

static ConvArgs base_args(const Ctx& c) {
    ConvArgs a{};
    a.B = c.P.B; a.G = c.P.G;
    return a;
}

static int launch_pw(const tcpw::Args& a, int G, cudaStream_t s, int cat) {
    ENDO_SET_MAX_SMEM(tcpw::pw_gemm_kernel, 227 * 1024);
    const size_t smem = tcpw::smem_bytes(a.Npad, a.K, a.mode);
    if (smem > 227 * 1024) return ENDO_ERR_CONFIG;
    dim3 grid(cdiv(a.per_group, a.mode == 2 ? tcpw::MT / 4 : tcpw::MT), G, 1);
    ProfScope prof(cat, s);
    launch_pdl(tcpw::pw_gemm_kernel, grid, tcpw::NTHREADS, smem, s, a);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

#define ENDO_TRY(expr)                 \
    do {                               \
        int _e = (expr);               \
        if (_e != ENDO_OK) return _e;  \
    } while (0)

// one launch packs the tensor-core weight images of every DenseLayer (forward / data gradient)
template <class F>
static void for_each_dense(const NetPlan& P, F f) {
    const int nd = P.cfg.n_down;
    for (int l = 0; l <= nd; ++l) for (const auto& d : P.down[l]) f(d);
    for (int i = 0; i < nd; ++i) for (const auto& d : P.up[i]) f(d);
}
// images: every DenseLayer + every 16-output-channel pass of the TransitionUp convolutions; tables of <= 112 entries per launch
template <class K>
static int pack_all(const Ctx& c, bool bwd, int per, int mode, K kern, unsigned char* region) {
    const NetPlan& P = c.P;
    tcconv::PackTable T{};
    T.mode = mode;
    auto flush = [&]() -> int {
        if (T.n == 0) return ENDO_OK;
        ProfScope prof(PC_BN, c.s);
        launch_pdl(kern, T.total_chunks, 256, 0, c.s, c.params, region, T);
        ENDO_CHECK_LAUNCH();
        T.n = 0; T.total_chunks = 0;
        return ENDO_OK;
    };
    int rc = ENDO_OK;
    auto add = [&](long long w, int K_, int N_, long long out) {
        if (rc != ENDO_OK) return;
        if (T.n == 112) rc = flush();
        T.e[T.n++] = tcconv::PackEntry{w, K_, N_, T.total_chunks, out};
        T.total_chunks += cdiv(K_, per);
    };
    for_each_dense(P, [&](const DenseLayerP& d) { add(d.conv.w, d.cin, d.conv.cout, bwd ? d.wpb_off : d.wp_off); });
    for (int i = 0; i < P.cfg.n_down; ++i) {
        const TransUpP& t = P.tu[i];
        if (t.conv.cout > 128) return ENDO_ERR_CONFIG;
        for (int q = 0; q * 16 < t.conv.cout; ++q) {
            const int n = (t.conv.cout - q * 16) < 16 ? (t.conv.cout - q * 16) : 16;
            add(t.conv.w + (long long)q * 16 * t.cin * 9, t.cin, n, bwd ? t.wpb_off[q] : t.wp_off[q]);
        }
    }
    if (!bwd && P.first.cout <= 128)                     // first convolution (forward on the persistent kernel, 16 output channels per pass)
        for (int q = 0; q * 16 < P.first.cout; ++q) {
            const int n = (P.first.cout - q * 16) < 16 ? (P.first.cout - q * 16) : 16;
            add(P.first.w + (long long)q * 16 * P.cfg.in_channels * 9, P.cfg.in_channels, n, P.first_wp_off[q]);
        }
    if (rc != ENDO_OK) return rc;
    return flush();
}
static int pack_dense_weights_fwd(const Ctx& c) {
    const int mode = x3_mode(c.math) == 3 ? 2 : x3_mode(c.math);       // plain bf16 uses the bf16x3 images (first terms)
    return pack_all(c, false, mode == 1 ? 8 : 16, mode, tcconv::pack_w_fwd_all_kernel, reinterpret_cast<unsigned char*>(c.acts + c.P.wpack_off));
}
static int pack_dense_weights_bwd(const Ctx& c) {
    return pack_all(c, true, 64, 0, tcconv::pack_w_dgrad_all_kernel, reinterpret_cast<unsigned char*>(c.scratch + c.P.wpack_bwd_off));
}

static int dense_layer_fwd(const Ctx& c, const DenseLayerP& d) {
    const NetPlan& P = c.P;
    const int l = d.level;
    ENDO_TRY(bn_prepare(c, d.bn, l, d.in_off));
    ConvArgs a = base_args(c);
    a.in = c.X(l); a.coef = c.COEF(d.bn); a.in_C = P.Ctot[l]; a.in_off = d.in_off; a.K = d.cin; a.ih = P.h[l]; a.iw = P.w[l];
    a.w = c.params + d.conv.w; a.bias = c.params + d.conv.b; a.w_cin = d.cin;
    a.out = c.X(l); a.out_C = P.Ctot[l]; a.out_off = d.out_off; a.N = d.conv.cout; a.oh = P.h[l]; a.ow = P.w[l];
    a.stats = c.ST(l); a.stats_C = P.Ctot[l];
    if (is_tc(c.math) && !(tc_disable_mask() & 1)) {
        // tcgen05 path: tf32 operands (what cuDNN runs the reference's convs in by default), fp32 accumulate in TMEM
        tcconv::FwdArgs t;
        t.in = a.in; t.coef = a.coef; t.w = a.w; t.bias = a.bias; t.out = a.out; t.stats = a.stats;
        t.in_C = a.in_C; t.in_off = a.in_off; t.K = a.K; t.out_C = a.out_C; t.out_off = a.out_off; t.N = a.N;
        t.H = a.oh; t.W = a.ow; t.B = a.B; t.G = a.G; t.stats_C = a.stats_C; t.up = 0; t.dbg = tc_debug_mask(); t.one = 0;
        t.x3 = x3_mode(c.math);
        t.wpack = reinterpret_cast<const float*>(c.acts + P.wpack_off + d.wp_off);   // packed by pack_dense_weights_fwd()
        // 3xTF32 at the high-resolution levels: the persistent TMA-fed kernel of net_fwd2.cuh (ENDO_TC_DISABLE bit 32768: round-1 kernel)
        {
            const int tx2 = cdiv(t.W, tcfwd2::TW), ty2 = cdiv(t.H, tcfwd2::TH);
            const int n2 = tx2 * ty2 * t.B;
            if (t.x3 == 1 && n2 >= 4 * kNumSMs && t.K <= tcfwd2::COEF_MAX && t.N <= tcfwd2::OUT_MAXN && (t.N & 3) == 0 &&
                !(tc_disable_mask() & 32768)) {
                tcfwd2::Args f;
                f.coef = t.coef; f.bias = t.bias; f.wpack = t.wpack; f.stats = t.stats;
                f.in_off = t.in_off; f.K = t.K; f.out_off = t.out_off; f.N = t.N; f.H = t.H; f.W = t.W; f.B = t.B; f.G = t.G;
                f.stats_C = t.stats_C; f.tiles_x = tx2; f.tiles_y = ty2; f.n_tiles = n2; f.up = 0; f.dbg = tc_debug_mask();
                CUtensorMap in_map, out_map;
                if (!tma::make_nhwc_map(&in_map, t.in, t.B, t.H, t.W, t.in_C, 8, tcfwd2::PITCH, tcfwd2::TH + 2) ||
                    !tma::make_nhwc_map(&out_map, t.out, t.B, t.H, t.W, t.out_C, t.N, tcfwd2::TW, tcfwd2::TH))
                    return ENDO_ERR_CUDA;
                ENDO_SET_MAX_SMEM(tcfwd2::dense_fwd_x3_persistent_kernel, tcfwd2::SMEM_BYTES);
                ProfScope prof(PC_CONV_DENSE_FWD, c.s);
                launch_pdl(tcfwd2::dense_fwd_x3_persistent_kernel, n2 < kNumSMs ? n2 : kNumSMs, tcfwd2::NTHREADS, tcfwd2::SMEM_BYTES, c.s, f, in_map, out_map);
                ENDO_CHECK_LAUNCH();
                return ENDO_OK;
            }
        }
        ENDO_SET_MAX_SMEM(tcconv::dense_fwd_tf32_kernel, tcconv::SMEM_BYTES);
        // Low-resolution levels: a CTA per 32x32 tile over ALL input channels leaves most SMs idle behind a long serial
        // channel loop.  Split the channel chunks over blockIdx.y (raw partial sums to scratch, splitk_finish_kernel adds
        // the slices in a fixed order, then bias + statistics): pick the slice count that minimises waves x chunks.
        const int tiles = cdiv(t.W, tcconv::TW) * cdiv(t.H, tcconv::TH) * t.B;
        const int nchunks = cdiv(t.K, t.x3 == 1 ? 8 : 16);
        const long long pixels = (long long)t.B * t.H * t.W;
        int ksplit = 1;
        if (tiles < 2 * kNumSMs && !(tc_disable_mask() & 128)) {
            long long best = -1;
            for (int ks = 1; ks <= 16 && ks <= nchunks; ++ks) {
                if (4ll * ks * pixels * 16 > P.tdtmp_bytes) break;
                const long long cost = (long long)cdiv((long long)tiles * ks, kNumSMs) * (cdiv(nchunks, ks) + 2);
                if (best < 0 || cost < best) { best = cost; ksplit = ks; }
            }
            ksplit = cdiv(nchunks, cdiv(nchunks, ksplit));          // no empty slice
        }
        t.partial = nullptr; t.ksplit = 1; t.pixels = pixels;
        if (ksplit > 1) { t.partial = reinterpret_cast<float*>(c.acts + P.tdtmp_off); t.ksplit = ksplit; }
        dim3 grid(cdiv(t.W, tcconv::TW) * cdiv(t.H, tcconv::TH), ksplit, t.B);
        ProfScope prof(PC_CONV_DENSE_FWD, c.s);
        launch_pdl(tcconv::dense_fwd_tf32_kernel, grid, tcconv::NTHREADS, tcconv::SMEM_BYTES, c.s, t);
        ENDO_CHECK_LAUNCH();
        if (ksplit > 1) {
            const int per_group = (int)(pixels / t.G);
            int fblocks = cdiv(per_group, 64);
            if (fblocks > 4 * kNumSMs) fblocks = 4 * kNumSMs;
            launch_pdl(splitk_finish_kernel<16>, dim3(fblocks, t.G), 256, 0, c.s, t.partial, t.bias, t.out, t.stats, ksplit, pixels, per_group,
                                                                        t.N, t.out_C, t.out_off, t.stats_C);
            ENDO_CHECK_LAUNCH();
        }
        return ENDO_OK;
    }
    // Low-resolution levels: 32x32 tiles over all input channels would occupy a handful of SMs for hundreds of
    // microseconds.  Use 8x32 tiles and split the input channels over blockIdx.y; a second tiny kernel adds the slices.
    const int big_tiles = cdiv(a.ow, 32) * cdiv(a.oh, 32) * a.B;
    if (big_tiles < 2 * kNumSMs && a.K >= 32 && !(tc_disable_mask() & 128)) {
        const int tiles = cdiv(a.ow, 32) * cdiv(a.oh, 8) * a.B;
        int ksplit = cdiv(3 * kNumSMs, tiles);
        if (ksplit > a.K / 16) ksplit = a.K / 16;
        if (ksplit < 1) ksplit = 1;
        const long long pixels = (long long)a.B * a.oh * a.ow;
        if (4ll * ksplit * pixels * 16 <= P.tdtmp_bytes) {
            a.partial = reinterpret_cast<float*>(c.acts + P.tdtmp_off);
            a.ksplit = ksplit;
            const int per_group = (int)(pixels / a.G);
            int fblocks = cdiv(per_group, 64);
            if (fblocks > 4 * kNumSMs) fblocks = 4 * kNumSMs;
            if (d.conv.cout == 12) {
                ENDO_TRY((launch_conv_splitk<3, 2, 12, 4, LM_BNRELU>(a, c.s)));
                ProfScope prof(PC_CONV_DENSE_FWD, c.s);
                launch_pdl(splitk_finish_kernel<12>, dim3(fblocks, a.G), 256, 0, c.s, a.partial, a.bias, a.out, a.stats, ksplit, pixels, per_group,
                                                                            a.N, a.out_C, a.out_off, a.stats_C);
            } else {
                ENDO_TRY((launch_conv_splitk<3, 2, 16, 4, LM_BNRELU>(a, c.s)));
                ProfScope prof(PC_CONV_DENSE_FWD, c.s);
                launch_pdl(splitk_finish_kernel<16>, dim3(fblocks, a.G), 256, 0, c.s, a.partial, a.bias, a.out, a.stats, ksplit, pixels, per_group,
                                                                            a.N, a.out_C, a.out_off, a.stats_C);
            }
            ENDO_CHECK_LAUNCH();
            return ENDO_OK;
        }
    }
    if (a.K <= 1536 && !(tc_disable_mask() & 256)) {
        if (d.conv.cout == 12) return launch_conv_pf<8, 12>(a, c.s);
        return launch_conv_pf<6, 16>(a, c.s);
    }
    if (d.conv.cout == 12) return launch_conv<3, 8, 12, 4, LM_BNRELU, EM_STORE, WM_FWD, false>(a, c.s);
    return launch_conv<3, 6, 16, 4, LM_BNRELU, EM_STORE, WM_FWD, false>(a, c.s);
}

// Weight-gradient launch (DenseLayer / TransitionUp / TransitionDown 1x1 modes): the TMA-fed kernel of net_wgrad2.cuh, one
// persistent CTA per SM (channel blocks x tile ranges); ENDO_TC_DISABLE bit 16384 selects the round-1 register-staged kernel.
static int launch_wgrad_tc(tcwgrad::Args t, int cin, cudaStream_t s, int cat) {
    const int yblocks = cdiv(cin, tcwgrad::MCH);
    t.n_tiles = t.B * cdiv(t.H, tcwgrad::TR) * cdiv(t.W, tcwgrad::TW);
    t.dbg = tc_debug_mask();
    if (tc_disable_mask() & 16384) {
        int want = (2 * kNumSMs) / yblocks;
        if (want < 1) want = 1;
        if (want > t.n_tiles) want = t.n_tiles;
        t.tiles_per_cta = cdiv(t.n_tiles, want);
        dim3 grid(cdiv(t.n_tiles, t.tiles_per_cta), yblocks, 1);
        ProfScope prof(cat, s);
        tcwgrad::dense_wgrad_bf16_kernel<<<grid, tcwgrad::NTHREADS, tcwgrad::SMEM_BYTES, s>>>(t);
        ENDO_CHECK_LAUNCH();
        return ENDO_OK;
    }
    t.n_tiles = t.B * cdiv(t.H, tcwgrad2::TR) * cdiv(t.W, tcwgrad2::TW);
    int want = kNumSMs / yblocks;                            // one wave of persistent CTAs
    if (want < 1) want = 1;
    if (want > t.n_tiles) want = t.n_tiles;
    t.tiles_per_cta = cdiv(t.n_tiles, want);
    const int sh = t.up ? 1 : 0;
    CUtensorMap xmap;
    if (!tma::make_nhwc_map(&xmap, t.xa, t.B, t.H >> sh, t.W >> sh, t.xa_C, tcwgrad::MCH, tcwgrad2::TW >> sh, tcwgrad2::TR >> sh))
        return ENDO_ERR_CUDA;
    ENDO_SET_MAX_SMEM(tcwgrad2::dense_wgrad_tma_kernel, 227 * 1024);
    const size_t smem = tcwgrad2::smem_bytes(t.Cout, t.one);
    if (smem > 227 * 1024) return ENDO_ERR_CONFIG;
    dim3 grid(cdiv(t.n_tiles, t.tiles_per_cta), yblocks, 1);
    ProfScope prof(cat, s);
    launch_pdl(tcwgrad2::dense_wgrad_tma_kernel, grid, tcwgrad2::NTHREADS, smem, s, t, xmap);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

// weight gradient of a 3x3 convolution as a TMA -> tcgen05 GEMM over plane-major bf16 operands (net_wgrad3.cuh), on the
// weight-gradient stream: act16 = [ceil(Cin/8)][B*H*W][8] activations, g16 = [g_groups][B*H*W][8] output gradient of which the
// two groups from g_grp0 (16 output channels) are used; dw = OIHW block of those output channels
static int launch_wgrad_gemm(const Ctx& c, float* dw, int Cin, int Cout, int H, int W, const unsigned short* a16, const unsigned short* g16,
                             int g_groups, int g_grp0, int cat, bool swap = false) {
    const NetPlan& P = c.P;
    // swap (Cin <= 8, Cout <= 128): the gradient planes are the M operand, ALL output channels in one pass (net_wgrad3.cuh)
    const int c8 = swap ? g_groups * 8 : (Cin + 7) / 8 * 8;      // channels of the M operand
    tcwgrad3::Args g{};
    g.dw = dw; g.Cin = Cin; g.Cout = Cout; g.H = H; g.W = W; g.g_grp0 = g_grp0; g.swap = swap ? 1 : 0;
    g.tiles_x = cdiv(W, tcwgrad3::TW); g.tiles_y = cdiv(H, tcwgrad3::TR); g.n_tiles = g.tiles_x * g.tiles_y * P.B;
    const int mb_all = cdiv(swap ? Cout : Cin, 128);
    int ny = 1;
    g.groups = c8 / 8; g.mblocks = mb_all;
    // few tiles, or two stages of all channel groups would not fit: one 128-channel block per CTA
    if (mb_all > 1 && (g.n_tiles < 2 * kNumSMs || tcwgrad3::smem_bytes(g.groups, g.mblocks, 2) > (size_t)tcwgrad3::SMEM_LIMIT)) {
        ny = mb_all; g.groups = 16; g.mblocks = 1;
    }
    g.sets = 512 / (g.mblocks * tcwgrad3::NB);
    if (g.sets > 3) g.sets = 3;
    g.nstages = 1;
    while (g.nstages < tcwgrad3::MAX_STAGES && tcwgrad3::smem_bytes(g.groups, g.mblocks, g.nstages + 1) <= (size_t)tcwgrad3::SMEM_LIMIT) ++g.nstages;
    const size_t smem = tcwgrad3::smem_bytes(g.groups, g.mblocks, g.nstages);
    if (smem > (size_t)tcwgrad3::SMEM_LIMIT || g.sets < 1) return ENDO_ERR_CONFIG;
    const int ctas = kNumSMs < g.n_tiles ? kNumSMs : g.n_tiles;
    g.tiles_per_cta = cdiv(g.n_tiles, ctas);
    CUtensorMap amap, gmap;
    if (swap) {
        if (!tcwgrad3::make_map(&amap, g16, P.B, H, W, tcwgrad3::TR, g_groups, g.groups) ||
            !tcwgrad3::make_map(&gmap, a16, P.B, H, W, tcwgrad3::TR, 1, 2))
            return ENDO_ERR_CUDA;
    } else if (!tcwgrad3::make_map(&amap, a16, P.B, H, W, tcwgrad3::TR, c8 / 8, g.groups) ||
               !tcwgrad3::make_map(&gmap, g16, P.B, H, W, tcwgrad3::TR, g_groups, 2))
        return ENDO_ERR_CUDA;
    ENDO_SET_MAX_SMEM(tcwgrad3::dense_wgrad_gemm_kernel, tcwgrad3::SMEM_LIMIT);
    ProfScope prof(cat, c.sw);
    launch_pdl(tcwgrad3::dense_wgrad_gemm_kernel, dim3(cdiv(g.n_tiles, g.tiles_per_cta), ny), tcwgrad3::NTHREADS, smem, c.sw, g, amap, gmap);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

static int dense_layer_bwd(const Ctx& c, const DenseLayerP& d) {
    const NetPlan& P = c.P;
    const int l = d.level;
    // weight / bias gradient
    WgradArgs w{};
    w.a_in = c.X(l); w.a_coef = c.COEF(d.bn); w.a_C = P.Ctot[l]; w.a_off = d.in_off; w.a_K = d.cin; w.a_h = P.h[l]; w.a_w = P.w[l];
    w.g_in = c.GX(l); w.g_x = c.X(l); w.g_ab = c.AB(l); w.g_C = P.Ctot[l]; w.g_off = d.out_off; w.g_K = d.conv.cout;
    w.g_h = P.h[l]; w.g_w = P.w[l]; w.oh = P.h[l]; w.ow = P.w[l]; w.B = P.B; w.G = P.G;
    w.dw = c.gparams + d.conv.w; w.db = c.gparams + d.conv.b; w.w_cin = d.cin;
    const bool tc_w = is_tc(c.math) && !(tc_disable_mask() & 4);
    const bool tc_d = is_tc(c.math) && !(tc_disable_mask() & 2);
    // weight gradient as a pure TMA -> MMA GEMM over the bf16 by-products of the data-gradient kernel (net_wgrad3.cuh); it runs
    // AFTER that kernel, on the main stream (one by-product buffer serves all layers); ENDO_TC_DISABLE bit 262144: round-2a kernel
    const bool gemm_w = tc_w && tc_d && d.cin <= 384 && d.conv.cout <= 16 && !(tc_disable_mask() & 262144);
    if (!gemm_w) ENDO_TRY(c.fork());
    // the conv bias gradient is produced by exactly one kernel: the tcgen05 dgrad if it runs, else the FFMA wgrad if it
    // runs, else a tiny dedicated reduction
    if (tc_d) w.db = nullptr;
    if (tc_w && !tc_d) {
        ProfScope prof(PC_WGRAD, c.sw);
        launch_pdl(bias_grad_kernel, kNumSMs, 256, 0, c.sw, c.GX(l), c.X(l), c.AB(l), c.gparams + d.conv.b, P.Ctot[l], d.out_off, d.conv.cout,
                                                  (long long)(P.B / P.G) * P.h[l] * P.w[l], P.G);
        ENDO_CHECK_LAUNCH();
    }
    if (gemm_w) {
    } else if (tc_w) {
        // bf16 tensor-core weight gradient (pixels are the GEMM K dimension); the bias gradient comes from the dgrad kernel
        tcwgrad::Args t;
        t.x = c.X(l); t.coef = c.COEF(d.bn); t.g = c.GX(l); t.ab = c.AB(l); t.dw = c.gparams + d.conv.w;
        t.xa = c.X(l); t.xa_C = P.Ctot[l]; t.up = 0; t.one = 0;
        t.C = P.Ctot[l]; t.in_off = d.in_off; t.Cin = d.cin; t.out_off = d.out_off; t.Cout = d.conv.cout;
        t.H = P.h[l]; t.W = P.w[l]; t.B = P.B; t.G = P.G;
        ENDO_TRY(launch_wgrad_tc(t, d.cin, c.sw, PC_WGRAD));
    } else if (d.conv.cout == 12) {
        ENDO_TRY((launch_wgrad2<3, 12, 1, 4, LM_BNRELU, LM_GRAD, false>(w, c.sw)));
    } else {
        ENDO_TRY((launch_wgrad2<3, 16, 1, 4, LM_BNRELU, LM_GRAD, false>(w, c.sw)));
    }
    // data gradient through conv, ReLU and BatchNorm (first term; the mean terms are applied lazily)
    ConvArgs a = base_args(c);
    a.in = c.GX(l); a.in2 = c.X(l); a.in_ab = c.AB(l); a.in_C = P.Ctot[l]; a.in_off = d.out_off; a.K = d.conv.cout;
    a.ih = P.h[l]; a.iw = P.w[l];
    a.w = c.params + d.conv.w; a.w_cin = d.cin;
    a.out = c.GX(l); a.out_C = P.Ctot[l]; a.out_off = d.in_off; a.N = d.cin; a.oh = P.h[l]; a.ow = P.w[l];
    a.stats = c.BNRED(); a.stats_C = P.maxC;
    a.x = c.X(l); a.ep_coef = c.COEF(d.bn);
    if (tc_d) {
        tcdgrad::Args t{};
        t.g = c.GX(l); t.x = c.X(l); t.ab = c.AB(l); t.coef = c.COEF(d.bn); t.w = c.params + d.conv.w;
        t.gout = c.GX(l); t.db = c.gparams + d.conv.b; t.red = c.BNRED(); t.red_C = P.maxC;
        t.C = P.Ctot[l]; t.out_off = d.out_off; t.Cout = d.conv.cout; t.in_off = d.in_off; t.Cin = d.cin;
        t.H = P.h[l]; t.W = P.w[l]; t.B = P.B; t.G = P.G;
        t.wpack = reinterpret_cast<const float*>(c.scratch + P.wpack_bwd_off + d.wpb_off);   // packed by pack_dense_weights_bwd()
        int bk = 0;
        if (gemm_w) {
            ENDO_TRY(c.acquire_buf(&bk));
            t.a16 = c.A16(bk); t.g16 = c.G16(bk);
        }
        ENDO_SET_MAX_SMEM(tcdgrad::dense_dgrad_tf32_kernel, tcdgrad::SMEM_BYTES);
        const int tiles = cdiv(t.W, tcconv::TW) * cdiv(t.H, tcconv::TH);
        const int ysplit = (tiles * t.B < 4 * kNumSMs && !(tc_disable_mask() & 128)) ? cdiv(t.Cin, tcdgrad::NC) : 1;
        dim3 grid(tiles, ysplit, t.B);
        {
            ProfScope prof(PC_DGRAD, c.s);
            launch_pdl(tcdgrad::dense_dgrad_tf32_kernel, grid, tcdgrad::NTHREADS, tcdgrad::SMEM_BYTES, c.s, t);
            ENDO_CHECK_LAUNCH();
        }
        if (gemm_w) {
            ENDO_TRY(c.fork());                              // by-products complete -> the GEMM overlaps the next layers' data gradients
            ENDO_TRY(launch_wgrad_gemm(c, c.gparams + d.conv.w, d.cin, d.conv.cout, t.H, t.W, t.a16, t.g16, 2, 0, PC_WGRAD));
            ENDO_TRY(c.release_buf(bk));
        }
    } else {
        // all 12 (16) output-gradient channels in ONE staging step (no padded K), 32 input channels per CTA so that two
        // CTAs fit an SM (<= 128 registers): their staging / epilogue phases overlap each other's FMA phase
        if (d.conv.cout == 12) ENDO_TRY((launch_conv<3, 2, 32, 8, LM_GRAD, EM_DGRAD_BN, WM_DGRAD, false, 12, 2>(a, c.s)));
        else ENDO_TRY((launch_conv<3, 2, 32, 8, LM_GRAD, EM_DGRAD_BN, WM_DGRAD, false, 16, 2>(a, c.s)));
    }
    BnBwdArgs b;
    b.red = c.BNRED(); b.red_C = P.maxC; b.coef = c.COEF(d.bn); b.mi = c.MI(l); b.ab = c.AB(l);
    b.dgamma = c.gparams + d.bn.gamma; b.dbeta = c.gparams + d.bn.beta;
    b.C = d.cin; b.Ctot = P.Ctot[l]; b.ch_off = d.in_off; b.G = P.G; b.count = c.count(l);
    ProfScope prof(PC_BN, c.s);
    launch_pdl(bn_bwd_finalize_kernel, cdiv(d.cin, 128), 128, 0, c.s, b);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

static int trans_down_fwd(const Ctx& c, int l) {
    const NetPlan& P = c.P;
    const TransDownP& t = P.td[l];
    const int cs = t.bn.c;
    ENDO_TRY(bn_prepare(c, t.bn, l, P.offIn[l]));
    ConvArgs a = base_args(c);
    a.in = c.X(l); a.coef = c.COEF(t.bn); a.in_C = P.Ctot[l]; a.in_off = P.offIn[l]; a.K = cs; a.ih = P.h[l]; a.iw = P.w[l];
    a.w = c.params + t.conv.w; a.bias = c.params + t.conv.b; a.w_cin = cs;
    a.out = c.X(l + 1); a.out_C = P.Ctot[l + 1]; a.out_off = P.offIn[l + 1]; a.N = cs; a.oh = P.h[l]; a.ow = P.w[l];
    a.stats = c.ST(l + 1); a.stats_C = P.Ctot[l + 1];
    a.argmax_out = reinterpret_cast<unsigned char*>(c.acts + t.argmax);
    if (is_tc(c.math) && !(tc_disable_mask() & 64) && cs <= 128 * tcconv::POOL_MAXQ) {
        // tcgen05: 1x1 convolution in passes of 48 output channels into a scratch tensor, then one HBM-bound pooling pass
        ENDO_SET_MAX_SMEM(tcconv::dense_fwd_tf32_kernel, tcconv::SMEM_BYTES);
        float* tmp = reinterpret_cast<float*>(c.acts + P.tdtmp_off);
        const int npad = (cs + 15) / 16 * 16;
        const bool pw = npad <= 512 && !(tc_disable_mask() & 512);
        const bool fused_pool = pw && !(a.oh & 1) && !(a.ow & 1) && !(cs & 3) && !(tc_disable_mask() & 65536);
        if (pw) {
            // one GEMM over all output channels per 128-pixel tile (net_pw.cuh): the input is read once
            tcpw::Args q{};
            q.in = a.in; q.in_C = a.in_C; q.in_off = a.in_off; q.coef = a.coef; q.wpack = c.WPACK(); q.bias = a.bias;
            q.out = tmp; q.out_C = cs; q.out_off = 0; q.K = cs; q.N = cs; q.Npad = npad;
            q.per_group = (long long)(P.B / P.G) * a.oh * a.ow; q.mode = 0; q.x3 = x3_mode(c.math) == 1 || x3_mode(c.math) == 2;   // plain tf32 / bf16 modes: tf32 operands
            if (fused_pool) {
                // max-pool, argmax and statistics in the GEMM epilogue: the full-resolution conv output never reaches HBM
                q.out = a.out; q.out_C = a.out_C; q.out_off = a.out_off; q.argmax_out = a.argmax_out; q.red = a.stats; q.red_C = a.stats_C;
                q.H = a.oh; q.W = a.ow; q.per_group = (long long)(P.B / P.G) * (a.oh / 2) * (a.ow / 2); q.mode = 2;
            }
            {
                ProfScope prof(PC_BN, c.s);
                launch_pdl(tcpw::pack_w_pw_kernel, cdiv(cs, q.x3 ? 8 : 16), 256, 0, c.s, a.w, cs, npad, q.x3, 0, c.WPACK());
                ENDO_CHECK_LAUNCH();
            }
            ENDO_TRY(launch_pw(q, P.G, c.s, PC_CONV_TRANS_FWD));
        }
        for (int co0 = 0; co0 < (pw ? 0 : cs); co0 += 48) {
            tcconv::FwdArgs f;
            f.in = a.in; f.coef = a.coef; f.w = a.w; f.bias = a.bias + co0; f.out = tmp; f.stats = nullptr;
            f.in_C = a.in_C; f.in_off = a.in_off; f.K = cs; f.out_C = cs; f.out_off = co0; f.N = (cs - co0) < 48 ? (cs - co0) : 48;
            f.H = a.oh; f.W = a.ow; f.B = a.B; f.G = a.G; f.stats_C = 0; f.up = 0; f.dbg = 0; f.one = 1; f.wpack = c.WPACK();
            f.x3 = x3_mode(c.math); f.partial = nullptr; f.ksplit = 1; f.pixels = 0;
            {
                ProfScope prof(PC_BN, c.s);
                if (f.x3 == 1) launch_pdl(tcconv::pack_w_1x1_x3_kernel, cdiv(cs, 8), 256, 0, c.s, a.w, cs, cs, co0, c.WPACK());
                else if (f.x3 >= 2) launch_pdl(tcconv::pack_w_1x1_b3_kernel, cdiv(cs, 16), 256, 0, c.s, a.w, cs, cs, co0, reinterpret_cast<uint32_t*>(c.WPACK()));
                else launch_pdl(tcconv::pack_w_1x1_kernel, cdiv(cs, 16), 256, 0, c.s, a.w, cs, cs, co0, c.WPACK());
                ENDO_CHECK_LAUNCH();
            }
            dim3 grid(cdiv(f.W, tcconv::TW) * cdiv(f.H, tcconv::TH), 1, f.B);
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcconv::dense_fwd_tf32_kernel, grid, tcconv::NTHREADS, tcconv::SMEM_BYTES, c.s, f);
            ENDO_CHECK_LAUNCH();
        }
        if (!fused_pool) {
            const long long pixels = (long long)(P.B / P.G) * (a.oh / 2) * (a.ow / 2);
            int blocks = (int)((pixels + 7) / 8);
            if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
            if (blocks < 1) blocks = 1;
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcconv::td_pool_kernel, dim3(blocks, P.G), 256, sizeof(float) * 16 * cs, c.s, tmp, a.out, a.argmax_out, a.stats, P.B, a.oh, a.ow, cs,
                                                                                          a.out_C, a.out_off, P.G, a.stats_C);
            ENDO_CHECK_LAUNCH();
        }
        return ENDO_OK;
    }
    return launch_conv<1, 2, 48, 8, LM_BNRELU, EM_POOL, WM_FWD, false>(a, c.s);
}

static int trans_down_bwd(const Ctx& c, int l) {
    const NetPlan& P = c.P;
    const TransDownP& t = P.td[l];
    const int cs = t.bn.c;
    const unsigned char* am = reinterpret_cast<const unsigned char*>(c.acts + t.argmax);
    WgradArgs w{};
    w.a_in = c.X(l); w.a_coef = c.COEF(t.bn); w.a_C = P.Ctot[l]; w.a_off = P.offIn[l]; w.a_K = cs; w.a_h = P.h[l]; w.a_w = P.w[l];
    w.g_in = c.GX(l + 1); w.g_x = c.X(l + 1); w.g_ab = c.AB(l + 1); w.g_argmax = am; w.g_C = P.Ctot[l + 1];
    w.g_off = P.offIn[l + 1]; w.g_K = cs; w.g_h = P.h[l + 1]; w.g_w = P.w[l + 1];
    w.oh = P.h[l]; w.ow = P.w[l]; w.B = P.B; w.G = P.G;
    w.dw = c.gparams + t.conv.w; w.db = c.gparams + t.conv.b; w.w_cin = cs;
    const int npad = (cs + 15) / 16 * 16;
    const bool dgrad_pw = is_tc(c.math) && npad <= 512 && !(tc_disable_mask() & 1024);
    // weight gradient as one C x C GEMM over the bf16 by-products of the data-gradient kernel (net_pwwgrad.cuh): runs AFTER it
    const bool wgrad_gemm = dgrad_pw && !(cs & 15) && !(tc_disable_mask() & 32) && !(tc_disable_mask() & 131072);
    if (!wgrad_gemm) ENDO_TRY(c.fork());
    if (wgrad_gemm) {
    } else if (is_tc(c.math) && !(tc_disable_mask() & 32)) {
        // tcgen05 (bf16): the weight-gradient kernel in 1x1 mode, 48 output channels per launch; bias gradient = sum of the
        // routed (= of the pooled) gradient, reduced over the coarse buffer
        {
            ProfScope prof(PC_WGRAD_TRANS, c.sw);
            launch_pdl(bias_grad_kernel, dim3(kNumSMs / 4, cdiv(cs, 16)), 256, 0, c.sw, c.GX(l + 1), c.X(l + 1), c.AB(l + 1), c.gparams + t.conv.b,
                                                                             P.Ctot[l + 1], P.offIn[l + 1], cs,
                                                                             (long long)(P.B / P.G) * P.h[l + 1] * P.w[l + 1], P.G);
            ENDO_CHECK_LAUNCH();
        }
        for (int co0 = 0; co0 < cs; co0 += 48) {
            tcwgrad::Args q;
            q.x = c.X(l); q.coef = c.COEF(t.bn); q.g = nullptr; q.ab = nullptr; q.dw = c.gparams + t.conv.w;
            q.xa = c.X(l); q.xa_C = P.Ctot[l]; q.up = 0;
            q.C = P.Ctot[l]; q.in_off = P.offIn[l]; q.Cin = cs; q.out_off = co0; q.Cout = cs;
            q.H = P.h[l]; q.W = P.w[l]; q.B = P.B; q.G = P.G;
            q.one = 1; q.argmax = am; q.gc = c.GX(l + 1); q.xc = c.X(l + 1); q.abc = c.AB(l + 1); q.cC = P.Ctot[l + 1];
            q.c_off = P.offIn[l + 1]; q.cH = P.h[l + 1]; q.cW = P.w[l + 1];
            ENDO_TRY(launch_wgrad_tc(q, cs, c.sw, PC_WGRAD_TRANS));
        }
    } else {
        ENDO_TRY((launch_wgrad2<1, 48, 1, 8, LM_BNRELU, LM_GRADPOOL, false>(w, c.sw)));
    }
    ConvArgs a = base_args(c);
    a.in = c.GX(l + 1); a.in2 = c.X(l + 1); a.in_ab = c.AB(l + 1); a.argmax = am; a.in_C = P.Ctot[l + 1];
    a.in_off = P.offIn[l + 1]; a.K = cs; a.ih = P.h[l + 1]; a.iw = P.w[l + 1];
    a.w = c.params + t.conv.w; a.w_cin = cs;
    a.out = c.GX(l); a.out_C = P.Ctot[l]; a.out_off = P.offIn[l]; a.N = cs; a.oh = P.h[l]; a.ow = P.w[l];
    a.stats = c.BNRED(); a.stats_C = P.maxC;
    a.x = c.X(l); a.ep_coef = c.COEF(t.bn);
    if (dgrad_pw) {
        // tcgen05 (tf32): routed-gradient GEMM over all input channels per 128-pixel tile, BN-backward epilogue (net_pw.cuh)
        tcpw::Args q{};
        q.argmax = am; q.gc = c.GX(l + 1); q.xc = c.X(l + 1); q.abc = c.AB(l + 1); q.cC = P.Ctot[l + 1]; q.c_off = P.offIn[l + 1];
        q.H = P.h[l]; q.W = P.w[l]; q.wpack = c.WPACK_BWD();
        q.out = c.GX(l); q.out_C = P.Ctot[l]; q.out_off = P.offIn[l]; q.x = c.X(l); q.ep_coef = c.COEF(t.bn);
        q.red = c.BNRED(); q.red_C = P.maxC; q.K = cs; q.N = cs; q.Npad = npad;
        q.per_group = (long long)(P.B / P.G) * P.h[l] * P.w[l]; q.mode = 1; q.x3 = 0;
        if (wgrad_gemm) {
            q.r16 = reinterpret_cast<unsigned short*>(c.scratch + t.r16);
            q.a16 = reinterpret_cast<unsigned short*>(c.scratch + t.a16);
        }
        {
            ProfScope prof(PC_BN, c.s);
            launch_pdl(tcpw::pack_w_pw_kernel, cdiv(cs, 16), 256, 0, c.s, t.conv.w + c.params, cs, npad, 0, 1, c.WPACK_BWD());
            ENDO_CHECK_LAUNCH();
        }
        ENDO_TRY(launch_pw(q, P.G, c.s, PC_DGRAD_TRANS));
    } else {
        ENDO_TRY((launch_conv<1, 2, 48, 8, LM_GRADPOOL, EM_DGRAD_BN, WM_DGRAD, false>(a, c.s)));
    }
    if (wgrad_gemm) {
        ENDO_TRY(c.fork());                                  // the by-products are complete: the GEMM overlaps what follows
        {
            ProfScope prof(PC_WGRAD_TRANS, c.sw);
            launch_pdl(bias_grad_kernel, dim3(kNumSMs / 4, cdiv(cs, 16)), 256, 0, c.sw, c.GX(l + 1), c.X(l + 1), c.AB(l + 1),
                       c.gparams + t.conv.b, P.Ctot[l + 1], P.offIn[l + 1], cs, (long long)(P.B / P.G) * P.h[l + 1] * P.w[l + 1], P.G);
            ENDO_CHECK_LAUNCH();
        }
        tcpww::Args g{};
        g.dw = c.gparams + t.conv.w; g.C = cs;
        const int mblocks = (cs + 127) / 128;
        int ny = 1;
        g.Nper = cs;
        while (mblocks * g.Nper > 512 || g.Nper > 256) { ++ny; g.Nper = (cdiv(cs, ny) + 15) / 16 * 16; }
        g.sets = 512 / (mblocks * g.Nper);
        if (g.sets > 4) g.sets = 4;
        g.KT = 128;
        if (tcpww::smem_bytes(cs, g.Nper, 128, 2) > (size_t)tcpww::SMEM_LIMIT) g.KT = 64;
        g.nstages = 1;
        while (g.nstages < tcpww::MAX_STAGES && tcpww::smem_bytes(cs, g.Nper, g.KT, g.nstages + 1) <= (size_t)tcpww::SMEM_LIMIT) ++g.nstages;
        const size_t smem = tcpww::smem_bytes(cs, g.Nper, g.KT, g.nstages);
        if (smem > (size_t)tcpww::SMEM_LIMIT) return ENDO_ERR_CONFIG;
        const long long pix = (long long)P.B * P.h[l] * P.w[l];
        g.n_tiles = cdiv(pix, g.KT);
        int ctas = kNumSMs / ny;
        if (ctas > g.n_tiles) ctas = g.n_tiles;
        if (ctas < 1) ctas = 1;
        g.tiles_per_cta = cdiv(g.n_tiles, ctas);
        CUtensorMap amap, rmap;
        if (!tcpww::make_planes_map(&amap, c.scratch + t.a16, pix, cs, g.KT, cs / 8) ||
            !tcpww::make_planes_map(&rmap, c.scratch + t.r16, pix, cs, g.KT, g.Nper / 8))
            return ENDO_ERR_CUDA;
        ENDO_SET_MAX_SMEM(tcpww::pw_wgrad_kernel, tcpww::SMEM_LIMIT);
        ProfScope prof(PC_WGRAD_TRANS, c.sw);
        launch_pdl(tcpww::pw_wgrad_kernel, dim3(cdiv(g.n_tiles, g.tiles_per_cta), ny), tcpww::NTHREADS, smem, c.sw, g, amap, rmap);
        ENDO_CHECK_LAUNCH();
    }
    BnBwdArgs b;
    b.red = c.BNRED(); b.red_C = P.maxC; b.coef = c.COEF(t.bn); b.mi = c.MI(l); b.ab = c.AB(l);
    b.dgamma = c.gparams + t.bn.gamma; b.dbeta = c.gparams + t.bn.beta;
    b.C = cs; b.Ctot = P.Ctot[l]; b.ch_off = P.offIn[l]; b.G = P.G; b.count = c.count(l);
    ProfScope prof(PC_BN, c.s);
    launch_pdl(bn_bwd_finalize_kernel, cdiv(cs, 128), 128, 0, c.s, b);
    ENDO_CHECK_LAUNCH();
    return ENDO_OK;
}

static int trans_up_fwd(const Ctx& c, int i) {
    const NetPlan& P = c.P;
    const TransUpP& t = P.tu[i];
    const int l = t.dst_level, ls = t.src_level;
    ConvArgs a = base_args(c);
    a.in = c.X(ls); a.in_C = P.Ctot[ls]; a.in_off = t.src_off; a.K = t.cin; a.ih = P.h[ls]; a.iw = P.w[ls];
    a.w = c.params + t.conv.w; a.bias = c.params + t.conv.b; a.w_cin = t.cin;
    a.out = c.X(l); a.out_C = P.Ctot[l]; a.out_off = 0; a.N = t.conv.cout; a.oh = P.h[l]; a.ow = P.w[l];
    a.stats = c.ST(l); a.stats_C = P.Ctot[l];
    if (is_tc(c.math) && !(tc_disable_mask() & 8)) {
        // tcgen05: the DenseLayer forward kernel with the upsampling loader, 16 output channels per pass
        ENDO_SET_MAX_SMEM(tcconv::dense_fwd_tf32_kernel, tcconv::SMEM_BYTES);
        const int tx2 = cdiv(a.ow, tcfwd2::TW), ty2 = cdiv(a.oh, tcfwd2::TH), n2 = tx2 * ty2 * a.B;
        const bool use2 = x3_mode(c.math) == 1 && n2 >= 4 * kNumSMs && !(a.oh & 1) && !(a.ow & 1) && !(tc_disable_mask() & 32768);
        for (int co0 = 0; co0 < (use2 ? t.conv.cout : 0); co0 += 16) {
            // persistent TMA-fed kernel (net_fwd2.cuh) with the upsampling transform, 16 output channels per pass
            tcfwd2::Args f;
            f.coef = nullptr; f.bias = a.bias + co0; f.stats = a.stats;
            f.wpack = reinterpret_cast<const float*>(c.acts + P.wpack_off + t.wp_off[co0 / 16]);
            f.in_off = a.in_off; f.K = a.K; f.out_off = a.out_off + co0; f.N = (t.conv.cout - co0) < 16 ? (t.conv.cout - co0) : 16;
            f.H = a.oh; f.W = a.ow; f.B = a.B; f.G = a.G; f.stats_C = a.stats_C; f.tiles_x = tx2; f.tiles_y = ty2; f.n_tiles = n2;
            f.up = 1; f.dbg = 0;
            if (f.N & 3) return ENDO_ERR_CONFIG;
            CUtensorMap in_map, out_map;
            if (!tma::make_nhwc_map(&in_map, a.in, a.B, a.oh >> 1, a.ow >> 1, a.in_C, 8, tcfwd2::UP_W, tcfwd2::UP_H) ||
                !tma::make_nhwc_map(&out_map, a.out, a.B, a.oh, a.ow, a.out_C, f.N, tcfwd2::TW, tcfwd2::TH))
                return ENDO_ERR_CUDA;
            ENDO_SET_MAX_SMEM(tcfwd2::dense_fwd_x3_persistent_kernel, tcfwd2::SMEM_BYTES);
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcfwd2::dense_fwd_x3_persistent_kernel, n2 < kNumSMs ? n2 : kNumSMs, tcfwd2::NTHREADS, tcfwd2::SMEM_BYTES, c.s, f, in_map, out_map);
            ENDO_CHECK_LAUNCH();
        }
        for (int co0 = 0; co0 < (use2 ? 0 : t.conv.cout); co0 += 16) {
            tcconv::FwdArgs f;
            f.in = a.in; f.coef = nullptr; f.w = a.w + (size_t)co0 * t.cin * 9; f.bias = a.bias + co0; f.out = a.out; f.stats = a.stats;
            f.in_C = a.in_C; f.in_off = a.in_off; f.K = a.K; f.out_C = a.out_C; f.out_off = a.out_off + co0;
            f.N = (t.conv.cout - co0) < 16 ? (t.conv.cout - co0) : 16;
            f.H = a.oh; f.W = a.ow; f.B = a.B; f.G = a.G; f.stats_C = a.stats_C; f.up = 1; f.dbg = 0; f.one = 0;
            f.x3 = x3_mode(c.math); f.partial = nullptr; f.ksplit = 1; f.pixels = 0;
            f.wpack = reinterpret_cast<const float*>(c.acts + P.wpack_off + t.wp_off[co0 / 16]);    // packed by pack_dense_weights_fwd()
            dim3 grid(cdiv(f.W, tcconv::TW) * cdiv(f.H, tcconv::TH), 1, f.B);
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcconv::dense_fwd_tf32_kernel, grid, tcconv::NTHREADS, tcconv::SMEM_BYTES, c.s, f);
            ENDO_CHECK_LAUNCH();
        }
        return ENDO_OK;
    }
    return launch_conv<3, 2, 48, 8, LM_PLAIN, EM_STORE, WM_FWD, true>(a, c.s);
}

static int trans_up_bwd(const Ctx& c, int i) {
    const NetPlan& P = c.P;
    const TransUpP& t = P.tu[i];
    const int l = t.dst_level, ls = t.src_level;
    WgradArgs w{};
    w.a_in = c.X(ls); w.a_C = P.Ctot[ls]; w.a_off = t.src_off; w.a_K = t.cin; w.a_h = P.h[ls]; w.a_w = P.w[ls];
    w.g_in = c.GX(l); w.g_x = c.X(l); w.g_ab = c.AB(l); w.g_C = P.Ctot[l]; w.g_off = 0; w.g_K = t.conv.cout;
    w.g_h = P.h[l]; w.g_w = P.w[l]; w.oh = P.h[l]; w.ow = P.w[l]; w.B = P.B; w.G = P.G;
    w.dw = c.gparams + t.conv.w; w.db = c.gparams + t.conv.b; w.w_cin = t.cin;
    const long long tmp_need = 4ll * P.B * P.h[l] * P.w[l] * t.cin;
    const bool dgrad_tc = is_tc(c.math) && !(tc_disable_mask() & 2048) && t.cin <= tcdgrad::NC && (t.cin & 3) == 0 &&
                          tmp_need <= P.tdtmp_bytes;
    // weight gradient as TMA -> tcgen05 GEMMs (net_wgrad3.cuh): the upsampled input is packed to bf16 planes by a small kernel, the
    // output gradient planes are by-products of the data-gradient passes below; ENDO_TC_DISABLE bit 262144: round-2a kernel
    const bool gemm_w = dgrad_tc && !(tc_disable_mask() & 16) && !(tc_disable_mask() & 262144) && !(t.cin & 7) && t.cin <= 384 &&
                        !(P.h[l] & 1) && !(P.w[l] & 1);
    int bk = 0;
    if (gemm_w) {
        ENDO_TRY(c.acquire_buf(&bk));
        const long long items = (long long)P.B * P.h[l] * P.w[l] * (t.cin / 8);
        int blocks = (int)((items + 255) / 256);
        if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
        ProfScope prof(PC_WGRAD_TRANS, c.s);
        launch_pdl(tcwgrad3::upsample_pack16_kernel, blocks, 256, 0, c.s, c.X(ls), P.Ctot[ls], t.src_off, t.cin, c.A16(bk), P.B, P.h[l], P.w[l]);
        ENDO_CHECK_LAUNCH();
    } else {
        ENDO_TRY(c.fork());
    }
    if (gemm_w) {
    } else if (is_tc(c.math) && !(tc_disable_mask() & 16)) {
        // tcgen05 (bf16): the DenseLayer weight-gradient kernel with the upsampling loader, 16 output channels per pass;
        // the bias gradient comes from the small dedicated reduction
        if (!dgrad_tc) {                                  // otherwise the data-gradient passes below produce the bias gradient
            ProfScope prof(PC_WGRAD_TRANS, c.sw);
            launch_pdl(bias_grad_kernel, dim3(kNumSMs / 2, cdiv(t.conv.cout, 16)), 256, 0, c.sw, c.GX(l), c.X(l), c.AB(l), c.gparams + t.conv.b, P.Ctot[l], 0,
                                                                                       t.conv.cout, (long long)(P.B / P.G) * P.h[l] * P.w[l], P.G);
            ENDO_CHECK_LAUNCH();
        }
        for (int co0 = 0; co0 < t.conv.cout; co0 += 16) {
            const int nco = (t.conv.cout - co0) < 16 ? (t.conv.cout - co0) : 16;
            tcwgrad::Args q;
            q.x = c.X(l); q.coef = nullptr; q.g = c.GX(l); q.ab = c.AB(l); q.dw = c.gparams + t.conv.w + (size_t)co0 * t.cin * 9;
            q.xa = c.X(ls); q.xa_C = P.Ctot[ls]; q.up = 1; q.one = 0;
            q.C = P.Ctot[l]; q.in_off = t.src_off; q.Cin = t.cin; q.out_off = co0; q.Cout = nco;
            q.H = P.h[l]; q.W = P.w[l]; q.B = P.B; q.G = P.G;
            ENDO_TRY(launch_wgrad_tc(q, t.cin, c.sw, PC_WGRAD_TRANS));
        }
    } else {
        ENDO_TRY((launch_wgrad<3, 12, 4, 1, LM_PLAIN, LM_GRAD, true>(w, c.sw)));
    }
    if (dgrad_tc) {
        // tcgen05 (tf32): the DenseLayer data-gradient kernel in plain mode, 16 output-gradient channels per pass, raw result
        // into the (now idle) forward scratch tensor at full resolution; then the 2x2 fold into the half-resolution buffer
        float* tmp = reinterpret_cast<float*>(c.acts + P.tdtmp_off);
        ENDO_SET_MAX_SMEM(tcdgrad::dense_dgrad_tf32_kernel, tcdgrad::SMEM_BYTES);
        for (int co0 = 0; co0 < t.conv.cout; co0 += 16) {
            tcdgrad::Args q{};
            q.g = c.GX(l); q.x = c.X(l); q.ab = c.AB(l); q.coef = nullptr; q.w = c.params + t.conv.w + (size_t)co0 * t.cin * 9;
            q.gout = nullptr; q.red = nullptr; q.red_C = 0;
            // conv bias gradient = sum of the (corrected) output gradient: a by-product of staging it; when the FFMA weight
            // gradient runs (debug toggle 16) that kernel produces it instead
            q.db = (is_tc(c.math) && !(tc_disable_mask() & 16)) ? c.gparams + t.conv.b + co0 : nullptr;
            q.C = P.Ctot[l]; q.out_off = co0; q.Cout = (t.conv.cout - co0) < 16 ? (t.conv.cout - co0) : 16; q.in_off = 0; q.Cin = t.cin;
            q.H = P.h[l]; q.W = P.w[l]; q.B = P.B; q.G = P.G;
            q.plain = 1; q.first = co0 == 0; q.oC = t.cin; q.o_off = 0; q.po = tmp;
            if (gemm_w) q.g16 = c.G16(bk) + (size_t)(co0 / 8) * P.B * P.h[l] * P.w[l] * 8;     // groups co0/8, co0/8 + 1 of the gradient planes
            q.wpack = reinterpret_cast<const float*>(c.scratch + P.wpack_bwd_off + t.wpb_off[co0 / 16]);   // packed by pack_dense_weights_bwd()
            dim3 grid(cdiv(q.W, tcconv::TW) * cdiv(q.H, tcconv::TH), 1, q.B);
            ProfScope prof(PC_DGRAD_TRANS, c.s);
            launch_pdl(tcdgrad::dense_dgrad_tf32_kernel, grid, tcdgrad::NTHREADS, tcdgrad::SMEM_BYTES, c.s, q);
            ENDO_CHECK_LAUNCH();
        }
        const long long items = (long long)P.B * P.h[ls] * P.w[ls] * (t.cin / 4);
        int blocks = (int)((items + 255) / 256);
        if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
        {
            ProfScope prof(PC_DGRAD_TRANS, c.s);
            launch_pdl(tcdgrad::up_sum_kernel, blocks, 256, 0, c.s, tmp, t.cin, c.GX(ls), P.Ctot[ls], t.src_off, P.B, P.h[ls], P.w[ls], t.cin);
            ENDO_CHECK_LAUNCH();
        }
        if (gemm_w) {
            ENDO_TRY(c.fork());
            const int g_groups = (t.conv.cout + 15) / 16 * 2;
            for (int co0 = 0; co0 < t.conv.cout; co0 += 16) {
                const int nco = (t.conv.cout - co0) < 16 ? (t.conv.cout - co0) : 16;
                ENDO_TRY(launch_wgrad_gemm(c, c.gparams + t.conv.w + (size_t)co0 * t.cin * 9, t.cin, nco, P.h[l], P.w[l], c.A16(bk), c.G16(bk),
                                           g_groups, co0 / 8, PC_WGRAD_TRANS));
            }
            ENDO_TRY(c.release_buf(bk));
        }
        return ENDO_OK;
    }
    ConvArgs a = base_args(c);
    a.in = c.GX(l); a.in2 = c.X(l); a.in_ab = c.AB(l); a.in_C = P.Ctot[l]; a.in_off = 0; a.K = t.conv.cout;
    a.ih = P.h[l]; a.iw = P.w[l];
    a.w = c.params + t.conv.w; a.w_cin = t.cin;
    a.out = c.GX(ls); a.out_C = P.Ctot[ls]; a.out_off = t.src_off; a.N = t.cin; a.oh = P.h[l]; a.ow = P.w[l];
    return launch_conv<3, 2, 48, 8, LM_GRAD, EM_DGRAD_UP, WM_DGRAD, false>(a, c.s);
}

}  // namespace endo

using namespace endo;

extern "C" long long endo_net_param_count(const endo_net_config* cfg) {
    NetPlan P;
    if (build_plan(cfg, 0, 0, 0, 1, P) != ENDO_OK) return -1;
    return P.n_params;
}
extern "C" long long endo_net_buffer_count(const endo_net_config* cfg) {
    NetPlan P;
    if (build_plan(cfg, 0, 0, 0, 1, P) != ENDO_OK) return -1;
    return P.n_buffers;
}
extern "C" size_t endo_net_activation_bytes(const endo_net_config* cfg, int B, int H, int W) {
    NetPlan P;
    // sized for the largest group count that divides B (the layout only grows with G)
    int G = 1;
    if (B > 0 && B % 2 == 0) G = 2;
    if (build_plan(cfg, B, H, W, G, P) != ENDO_OK) return 0;
    return (size_t)P.acts_bytes;
}
extern "C" size_t endo_net_backward_scratch_bytes(const endo_net_config* cfg, int B, int H, int W) {
    NetPlan P;
    int G = 1;
    if (B > 0 && B % 2 == 0) G = 2;
    if (build_plan(cfg, B, H, W, G, P) != ENDO_OK) return 0;
    return (size_t)P.scratch_bytes;
}

extern "C" int endo_net_fwd(const endo_net_config* cfg, const float* x, const float* params, float* bn_buffers,
                            float* y, void* acts, size_t acts_bytes, int B, int H, int W, int groups, int training,
                            int math_and_flags, endo_stream_t stream) {
    const int math = math_and_flags & ENDO_MATH_MASK;
    if (math != ENDO_MATH_FP32 && !is_tc(math)) return ENDO_ERR_CONFIG;
    if (groups != 1 && groups != 2) return ENDO_ERR_CONFIG;
    if (!x || !params || !bn_buffers || !y || !acts) return ENDO_ERR_BAD_POINTER;
    if (!aligned16(x) || !aligned16(params) || !aligned16(y) || (reinterpret_cast<uintptr_t>(acts) & 255u))
        return ENDO_ERR_BAD_POINTER;
    if (B <= 0 || B > 65535) return ENDO_ERR_BAD_SHAPE;
    NetPlan P;
    ENDO_TRY(build_plan(cfg, B, H, W, groups, P));
    if (acts_bytes < (size_t)P.acts_bytes) return ENDO_ERR_WORKSPACE;
    Ctx c{P, static_cast<char*>(acts), nullptr, params, nullptr, bn_buffers, (cudaStream_t)stream, training, math};
    c.sw = c.s;
    const int nd = cfg->n_down;
    // statistics accumulate with atomics: clear them
    ENDO_CUDA(cudaMemsetAsync(c.acts + P.stat_off[0], 0, (size_t)(P.mi_off[0] - P.stat_off[0]), c.s));
    if (is_tc(math)) ENDO_TRY(pack_dense_weights_fwd(c));
    const int tx2 = cdiv(W, tcfwd2::TW), ty2 = cdiv(H, tcfwd2::TH), n2 = tx2 * ty2 * B;
    if (x3_mode(math) == 1 && cfg->in_channels <= 8 && P.first.cout <= 128 && !(P.first.cout & 3) && n2 >= 4 * kNumSMs &&
        32ll * B * H * W <= P.tdtmp_bytes && !(tc_disable_mask() & 524288)) {
        // firstconv on the persistent 3xTF32 kernel (net_fwd2.cuh), 16 output channels per pass: the NCHW images are repacked once
        // to NHWC with 8 channels (into the TransitionDown scratch, idle until the backward) so that TMA boxes can fetch them;
        // ENDO_TC_DISABLE bit 524288: the FFMA kernel below (0.45 ms instead of 0.16 ms at 16 x 256x320)
        float* x8 = reinterpret_cast<float*>(c.acts + P.tdtmp_off);
        {
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcfwd2::nchw_to_nhwc8_kernel, 8 * kNumSMs, 256, 0, c.s, x, x8, B, cfg->in_channels, (long long)H * W);
            ENDO_CHECK_LAUNCH();
        }
        for (int co0 = 0; co0 < P.first.cout; co0 += 16) {
            tcfwd2::Args f;
            f.coef = nullptr; f.bias = params + P.first.b + co0; f.stats = c.ST(0);
            f.wpack = reinterpret_cast<const float*>(c.acts + P.wpack_off + P.first_wp_off[co0 / 16]);
            f.in_off = 0; f.K = 8; f.out_off = P.offIn[0] + co0; f.N = (P.first.cout - co0) < 16 ? (P.first.cout - co0) : 16;
            f.H = H; f.W = W; f.B = B; f.G = P.G; f.stats_C = P.Ctot[0]; f.tiles_x = tx2; f.tiles_y = ty2; f.n_tiles = n2;
            f.up = 0; f.dbg = 0;
            CUtensorMap in_map, out_map;
            if (!tma::make_nhwc_map(&in_map, x8, B, H, W, 8, 8, tcfwd2::PITCH, tcfwd2::TH + 2) ||
                !tma::make_nhwc_map(&out_map, c.X(0), B, H, W, P.Ctot[0], f.N, tcfwd2::TW, tcfwd2::TH))
                return ENDO_ERR_CUDA;
            ENDO_SET_MAX_SMEM(tcfwd2::dense_fwd_x3_persistent_kernel, tcfwd2::SMEM_BYTES);
            ProfScope prof(PC_CONV_TRANS_FWD, c.s);
            launch_pdl(tcfwd2::dense_fwd_x3_persistent_kernel, n2 < kNumSMs ? n2 : kNumSMs, tcfwd2::NTHREADS, tcfwd2::SMEM_BYTES, c.s, f, in_map, out_map);
            ENDO_CHECK_LAUNCH();
        }
    } else {   // firstconv 3x3 in_channels -> first_conv_channels on the NCHW input (models.py:111-113, :172)
        ConvArgs a = base_args(c);
        a.in = x; a.K = cfg->in_channels; a.ih = H; a.iw = W;
        a.w = params + P.first.w; a.bias = params + P.first.b; a.w_cin = cfg->in_channels;
        a.out = c.X(0); a.out_C = P.Ctot[0]; a.out_off = P.offIn[0]; a.N = P.first.cout; a.oh = H; a.ow = W;
        a.stats = c.ST(0); a.stats_C = P.Ctot[0];
        ENDO_TRY((launch_conv<3, 2, 48, 8, LM_NCHW, EM_STORE, WM_FWD, false>(a, c.s)));
    }
    for (int l = 0; l < nd; ++l) {                           // models.py:175-178
        for (const auto& d : P.down[l]) ENDO_TRY(dense_layer_fwd(c, d));
        ENDO_TRY(trans_down_fwd(c, l));
    }
    for (const auto& d : P.down[nd]) ENDO_TRY(dense_layer_fwd(c, d));   // bottleneck, :180
    for (int i = 0; i < nd; ++i) {                           // :181-184
        ENDO_TRY(trans_up_fwd(c, i));
        for (const auto& d : P.up[i]) ENDO_TRY(dense_layer_fwd(c, d));
    }
    {   // abs(finalConv(out)), :186
        const long long npix = (long long)B * H * W;
        float* pre = reinterpret_cast<float*>(c.acts + P.pre_off);
        ProfScope prof(PC_FINAL, c.s);
        launch_pdl(final_fwd_kernel, cdiv(npix * 8, 256), 256, 0, c.s, c.X(0), params + P.final_.w, params + P.final_.b, pre, y,
                                                               npix, P.Ctot[0]);
        ENDO_CHECK_LAUNCH();
    }
    return ENDO_OK;
}

extern "C" int endo_net_bwd(const endo_net_config* cfg, const float* g_y, const float* x, const float* params,
                            float* g_params, float* g_x, void* acts, size_t acts_bytes, void* scratch,
                            size_t scratch_bytes, int B, int H, int W, int groups, int accumulate, int math_and_flags,
                            endo_stream_t stream) {
    const int math = math_and_flags & ENDO_MATH_MASK, flags = math_and_flags & ~ENDO_MATH_MASK;
    if (math != ENDO_MATH_FP32 && !is_tc(math)) return ENDO_ERR_CONFIG;
    if (groups != 1 && groups != 2) return ENDO_ERR_CONFIG;
    if (g_x != nullptr) return ENDO_ERR_CONFIG;              // train.py never differentiates w.r.t. the images
    if (!g_y || !x || !params || !g_params || !acts || !scratch) return ENDO_ERR_BAD_POINTER;
    if (!aligned16(g_y) || !aligned16(params) || !aligned16(g_params) || (reinterpret_cast<uintptr_t>(acts) & 255u) ||
        (reinterpret_cast<uintptr_t>(scratch) & 255u))
        return ENDO_ERR_BAD_POINTER;
    if (B <= 0 || B > 65535) return ENDO_ERR_BAD_SHAPE;
    NetPlan P;
    ENDO_TRY(build_plan(cfg, B, H, W, groups, P));
    if (acts_bytes < (size_t)P.acts_bytes || scratch_bytes < (size_t)P.scratch_bytes) return ENDO_ERR_WORKSPACE;
    if (P.Ctot[0] > 384) return ENDO_ERR_CONFIG;
    Ctx c{P, static_cast<char*>(acts), static_cast<char*>(scratch), params, g_params, nullptr, (cudaStream_t)stream, 1, math};
    SideLease lease;                                         // joined + returned to the pool on every exit path
    if (!(flags & ENDO_NET_SINGLE_STREAM) && !(tc_disable_mask() & 8192)) {
        int dev = 0;
        ENDO_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64) return ENDO_ERR_NO_DEVICE;
        lease.t = side_acquire(dev); lease.dev = dev; lease.caller = c.s;
    }
    c.sw = lease.t ? lease.t->s : c.s;
    c.ev_fork = lease.t ? lease.t->fork : nullptr;
    if (lease.t) { c.ev_buf[0] = lease.t->buf[0]; c.ev_buf[1] = lease.t->buf[1]; }
    c.lease = &lease;
    const int nd = cfg->n_down;
    // gradient buffers of levels >= 1, the lazy-correction arrays and the BN sums start at zero; the level-0
    // gradient buffer (the largest) is fully written by the finalConv backward and needs no clearing
    ENDO_CUDA(cudaMemsetAsync(c.scratch + P.gx_off[1], 0, (size_t)(P.scratch_bytes - P.gx_off[1]), c.s));
    if (!accumulate) ENDO_CUDA(cudaMemsetAsync(g_params, 0, sizeof(float) * (size_t)P.n_params, c.s));
    {
        const long long npix = (long long)B * H * W;
        const float* pre = reinterpret_cast<const float*>(c.acts + P.pre_off);
        const int ppc = (int)((npix + 2 * kNumSMs - 1) / (2 * kNumSMs));
        ProfScope prof(PC_FINAL, c.s);
        launch_pdl(final_bwd_kernel, cdiv(npix, ppc), 256, sizeof(float) * P.Ctot[0], c.s, g_y, pre, c.X(0), params + P.final_.w, c.GX(0), g_params + P.final_.w, g_params + P.final_.b, npix, P.Ctot[0], ppc);
        ENDO_CHECK_LAUNCH();
    }
    if (is_tc(math)) ENDO_TRY(pack_dense_weights_bwd(c));
    for (int i = nd - 1; i >= 0; --i) {
        for (int j = (int)P.up[i].size() - 1; j >= 0; --j) ENDO_TRY(dense_layer_bwd(c, P.up[i][j]));
        ENDO_TRY(trans_up_bwd(c, i));
    }
    for (int j = (int)P.down[nd].size() - 1; j >= 0; --j) ENDO_TRY(dense_layer_bwd(c, P.down[nd][j]));
    for (int l = nd - 1; l >= 0; --l) {
        ENDO_TRY(trans_down_bwd(c, l));
        for (int j = (int)P.down[l].size() - 1; j >= 0; --j) ENDO_TRY(dense_layer_bwd(c, P.down[l][j]));
    }
    {   // firstconv: weight gradient only
        WgradArgs w{};
        w.a_in = x; w.a_K = cfg->in_channels; w.a_h = H; w.a_w = W;
        w.g_in = c.GX(0); w.g_x = c.X(0); w.g_ab = c.AB(0); w.g_C = P.Ctot[0]; w.g_off = P.offIn[0]; w.g_K = P.first.cout;
        w.g_h = H; w.g_w = W; w.oh = H; w.ow = W; w.B = B; w.G = P.G;
        w.dw = g_params + P.first.w; w.db = g_params + P.first.b; w.w_cin = cfg->in_channels;
        ENDO_TRY(c.fork());
        if (is_tc(c.math) && cfg->in_channels <= 8 && !(tc_disable_mask() & 4) && !(tc_disable_mask() & 262144)) {
            // TMA -> tcgen05 GEMMs (net_wgrad3.cuh), 16 output channels per pass, over bf16 planes packed by two small kernels (all on the
            // weight-gradient stream; the second one also yields the bias gradient).  ENDO_TC_DISABLE bit 262144: the FFMA kernel below
            int bk = 0;
            ENDO_TRY(c.acquire_buf(&bk));
            const long long npix = (long long)B * H * W;
            {
                ProfScope prof(PC_WGRAD_TRANS, c.sw);
                launch_pdl(tcwgrad3::image_pack16_kernel, 8 * kNumSMs, 256, 0, c.sw, x, c.A16(bk), B, cfg->in_channels, (long long)H * W);
                ENDO_CHECK_LAUNCH();
                launch_pdl(tcwgrad3::grad_pack16_kernel, dim3(2 * kNumSMs, cdiv(P.first.cout, 16)), 256, 0, c.sw, c.GX(0), c.X(0), c.AB(0), c.G16(bk),
                           g_params + P.first.b, P.Ctot[0], P.offIn[0], P.first.cout, npix / P.G, P.G);
                ENDO_CHECK_LAUNCH();
            }
            const int g_groups = (P.first.cout + 15) / 16 * 2;
            if (P.first.cout <= 128) {                       // operands swapped: all output channels in one pass (this tail is not overlapped)
                ENDO_TRY(launch_wgrad_gemm(c, g_params + P.first.w, cfg->in_channels, P.first.cout, H, W, c.A16(bk), c.G16(bk), g_groups, 0,
                                           PC_WGRAD_TRANS, true));
            } else
            for (int co0 = 0; co0 < P.first.cout; co0 += 16) {
                const int nco = (P.first.cout - co0) < 16 ? (P.first.cout - co0) : 16;
                ENDO_TRY(launch_wgrad_gemm(c, g_params + P.first.w + (size_t)co0 * cfg->in_channels * 9, cfg->in_channels, nco, H, W, c.A16(bk),
                                           c.G16(bk), g_groups, co0 / 8, PC_WGRAD_TRANS));
            }
            ENDO_TRY(c.release_buf(bk));
        } else if (cfg->in_channels == 3 && P.first.cout == 48 && !(tc_disable_mask() & 4096)) {
            ProfScope prof(PC_WGRAD_TRANS, c.sw);
            launch_pdl(first_wgrad_kernel, 4 * kNumSMs, FW_THREADS, 0, c.sw, w);
            ENDO_CHECK_LAUNCH();
        } else {
            ENDO_TRY((launch_wgrad<3, 12, 4, 1, LM_NCHW, LM_GRAD, false>(w, c.sw)));
        }
    }
    return ENDO_OK;                                          // ~SideLease: the caller's stream waits for the weight gradients
}

extern "C" int endo_debug_trace_read(long long* host_out, int n) {
    if (!host_out || n <= 0 || n > 2048) return ENDO_ERR_BAD_SHAPE;
    ENDO_CUDA(cudaDeviceSynchronize());
    ENDO_CUDA(cudaMemcpyFromSymbol(host_out, endo::g_tc_trace, sizeof(long long) * (size_t)n));
    return ENDO_OK;
}
