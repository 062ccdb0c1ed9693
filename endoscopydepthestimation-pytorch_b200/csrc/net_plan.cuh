// Host-side plan of an FC-DenseNet (reference models.py:100-187): where every activation, statistic,
// parameter and gradient lives.  Shared by net_fwd.cu and net_bwd.cu.
//
// HBM layout (B200-first, not the reference's):
//   * activations are NHWC fp32;
//   * each resolution level owns ONE buffer [B, h, w, Ctot] that holds, side by side,
//         [ up | in | down-new | up-new ]
//       up       : output of the TransitionUp that lands on this level          (U  channels)
//       in       : input of the down DenseBlock (firstconv / TransitionDown out) (C0 channels)
//       down-new : the growth channels the down DenseBlock appends              (Dn channels)
//       up-new   : the growth channels the up DenseBlock appends                (Un channels)
//     so that `torch.cat([x, out], 1)` (models.py:46,52), the skip connection and
//     `torch.cat([up, skip], 1)` (models.py:79) are all zero-copy views: [in|down-new] is the skip,
//     [up|in|down-new] is the up block's input, and layers append in place.  The reference moves
//     495 MB per image through torch.cat at 256x320; here that traffic is zero.
//   * per-channel batch statistics (sum, sum of squares, fp64) are produced once per channel by the
//     kernel that writes the channel and shared by every later BatchNorm that reads it (the batch
//     mean/var of a channel does not depend on which BN module consumes it).
#pragma once
#include <vector>
#include "common.cuh"

namespace endo {

constexpr int kMaxLevels = 9;   // n_down <= 8 plus the bottleneck level



struct ConvP { long long w, b; int cin, cout, ks; };           // offsets (floats) into the flat parameter array
struct BnP { long long gamma, beta; long long rmean, rvar; int c; long long coef; };   // coef: float offset in acts
struct DenseLayerP { BnP bn; ConvP conv; int level, in_off, cin, out_off;
                     long long wp_off, wpb_off; };   // byte offsets of this layer's tensor-core weight images inside the wpack / wpack_bwd regions
struct TransDownP { BnP bn; ConvP conv; int level; long long argmax;                    // argmax: byte offset in acts
                    long long r16, a16; };      // byte offsets in the backward scratch: bf16 [B*h*w][C] routed gradient / relu(bn(x)),
                                                // by-products of the data-gradient kernel, operands of the weight-gradient GEMM
struct TransUpP { ConvP conv; int src_level, src_off, cin, dst_level;
                  long long wp_off[8], wpb_off[8]; };   // weight images of the 16-output-channel passes (cout <= 128)

struct NetPlan {
    endo_net_config cfg;
    int B, H, W, G, nlev;                       // nlev = n_down + 1
    int h[kMaxLevels], w[kMaxLevels];
    int U[kMaxLevels], C0[kMaxLevels], Dn[kMaxLevels], Un[kMaxLevels], Ctot[kMaxLevels];
    int offIn[kMaxLevels], offDn[kMaxLevels], offUn[kMaxLevels];
    // byte offsets inside the activation block
    long long x_off[kMaxLevels];                // float buffers [B,h,w,Ctot]
    long long stat_off[kMaxLevels];             // double [G][Ctot][2]  (sum, sumsq)
    long long mi_off[kMaxLevels];               // float  [G][Ctot][2]  (mean, invstd)
    long long pre_off;                          // float  [B*H*W] finalConv output before abs
    long long first_wp_off[8];                  // weight images of the first convolution's 16-output-channel passes (tensor-core forward)
    long long wpack_off;                        // sized for the widest layer: tensor-core weight image of the layer being run (forward)
    long long tdtmp_off, tdtmp_bytes;           // forward scratch: float [B,h,w,Cs] TransitionDown conv output before pooling (tensor-core
                                                // path, largest level); split-K partial sums of low-resolution DenseLayers
    long long acts_bytes;
    // byte offsets inside the backward scratch block
    long long gx_off[kMaxLevels];               // float [B,h,w,Ctot] gradient buffers
    long long ab_off[kMaxLevels];               // float [G][Ctot][2] lazy BN-backward correction (A, Bc)
    long long bnred_off;                        // double [G][maxC][2] per-layer BN backward sums
    long long a16_off[2], g16_off[2];           // (two sets: the GEMM of layer i reads one while the data gradient of layer i-1 writes
                                                // the other) bf16 by-products of the DenseLayer data gradient (operands of its weight-gradient GEMM):
                                                // relu(bn(x)) [B*h*w][cin rounded to 8] and the corrected output gradient [B*h*w][16]
    long long wpack_bwd_off;                    // sized for the widest layer: tensor-core weight image of the layer being run (backward)
    long long scratch_bytes;
    int maxC;
    ConvP first, final_;
    std::vector<DenseLayerP> down[kMaxLevels];  // down[l]: layers of denseBlocksDown.l ; down[n_down] = bottleneck
    std::vector<DenseLayerP> up[kMaxLevels];    // up[i]: layers of denseBlocksUp.i
    TransDownP td[kMaxLevels];
    TransUpP tu[kMaxLevels];
    long long n_params, n_buffers;
};

static inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// returns ENDO_OK or an error; B/H/W may be 0 when only the parameter counts are needed
static inline int build_plan(const endo_net_config* c, int B, int H, int W, int G, NetPlan& P) {
    if (!c) return ENDO_ERR_BAD_POINTER;
    if (c->n_down < 1 || c->n_down > 8 || c->growth_rate <= 0 || c->growth_rate % 4 || c->first_conv_channels <= 0 ||
        c->first_conv_channels % 4 || c->in_channels <= 0 || c->n_classes != 1 || c->bottleneck_layers < 1)
        return ENDO_ERR_CONFIG;
    if (c->growth_rate != 12 && c->growth_rate != 16) return ENDO_ERR_CONFIG;
    for (int i = 0; i < c->n_down; ++i)
        if (c->down_layers[i] < 1 || c->up_layers[i] < 1) return ENDO_ERR_CONFIG;
    P.cfg = *c;
    P.B = B; P.H = H; P.W = W; P.G = G < 1 ? 1 : G;
    const int nd = c->n_down, g = c->growth_rate;
    P.nlev = nd + 1;
    if (B > 0) {
        if (H <= 0 || W <= 0 || (H % (1 << nd)) || (W % (1 << nd))) return ENDO_ERR_BAD_SHAPE;
        if (B % P.G) return ENDO_ERR_BAD_SHAPE;
    }
    // ---- channel bookkeeping (models.py:114-163)
    int cur = c->first_conv_channels;
    for (int l = 0; l < nd; ++l) {
        P.C0[l] = cur; P.Dn[l] = g * c->down_layers[l];
        cur += P.Dn[l];
    }
    P.C0[nd] = cur; P.Dn[nd] = g * c->bottleneck_layers; P.U[nd] = 0; P.Un[nd] = 0;
    int prev = g * c->bottleneck_layers;
    for (int i = 0; i < nd; ++i) {
        const int l = nd - 1 - i;
        P.U[l] = prev; P.Un[l] = g * c->up_layers[i];
        prev = P.Un[l];
    }
    for (int l = 0; l <= nd; ++l) {
        P.offIn[l] = P.U[l];
        P.offDn[l] = P.U[l] + P.C0[l];
        P.offUn[l] = P.offDn[l] + P.Dn[l];
        P.Ctot[l] = P.offUn[l] + P.Un[l];
        P.h[l] = H >> l; P.w[l] = W >> l;
    }
    // ---- parameters, in state_dict() order
    long long po = 0, bo = 0;
    auto conv = [&](int cin, int cout, int ks) {
        ConvP q; q.cin = cin; q.cout = cout; q.ks = ks; q.w = po; po += (long long)cout * cin * ks * ks; q.b = po; po += cout;
        return q;
    };
    auto bn = [&](int ch) {
        BnP q; q.c = ch; q.gamma = po; po += ch; q.beta = po; po += ch; q.rmean = 0; q.rvar = 0; q.coef = 0;
        return q;
    };
    P.first = conv(c->in_channels, c->first_conv_channels, 3);
    for (int l = 0; l < nd; ++l) {
        P.down[l].clear();
        for (int j = 0; j < c->down_layers[l]; ++j) {
            DenseLayerP d; d.level = l; d.in_off = P.offIn[l]; d.cin = P.C0[l] + j * g; d.out_off = P.offDn[l] + j * g;
            d.bn = bn(d.cin); d.conv = conv(d.cin, g, 3);
            P.down[l].push_back(d);
        }
    }
    for (int l = 0; l < nd; ++l) {
        const int cs = P.C0[l] + P.Dn[l];
        P.td[l].level = l; P.td[l].bn = bn(cs); P.td[l].conv = conv(cs, cs, 1);
    }
    P.down[nd].clear();
    for (int j = 0; j < c->bottleneck_layers; ++j) {
        DenseLayerP d; d.level = nd; d.in_off = P.offIn[nd]; d.cin = P.C0[nd] + j * g; d.out_off = P.offDn[nd] + j * g;
        d.bn = bn(d.cin); d.conv = conv(d.cin, g, 3);
        P.down[nd].push_back(d);
    }
    for (int i = 0; i < nd; ++i) {
        const int l = nd - 1 - i;
        TransUpP& t = P.tu[i];
        t.dst_level = l; t.src_level = l + 1; t.cin = P.U[l];
        t.src_off = (i == 0) ? P.offDn[nd] : P.offUn[l + 1];
        t.conv = conv(P.U[l], P.U[l], 3);
    }
    for (int i = 0; i < nd; ++i) {
        const int l = nd - 1 - i;
        P.up[i].clear();
        for (int j = 0; j < c->up_layers[i]; ++j) {
            DenseLayerP d; d.level = l; d.in_off = 0; d.cin = P.offUn[l] + j * g; d.out_off = P.offUn[l] + j * g;
            d.bn = bn(d.cin); d.conv = conv(d.cin, g, 3);
            P.up[i].push_back(d);
        }
    }
    P.final_ = conv(P.Ctot[0], c->n_classes, 1);
    P.n_params = po;
    // ---- BN running buffers, in module order (running_mean[C], running_var[C] per BN)
    auto bnbuf = [&](BnP& q) { q.rmean = bo; bo += q.c; q.rvar = bo; bo += q.c; };
    for (int l = 0; l < nd; ++l) for (auto& d : P.down[l]) bnbuf(d.bn);
    for (int l = 0; l < nd; ++l) bnbuf(P.td[l].bn);
    for (auto& d : P.down[nd]) bnbuf(d.bn);
    for (int i = 0; i < nd; ++i) for (auto& d : P.up[i]) bnbuf(d.bn);
    P.n_buffers = bo;
    if (B <= 0) { P.acts_bytes = 0; P.scratch_bytes = 0; return ENDO_OK; }
    // ---- activation block
    long long off = 0;
    P.maxC = 0;
    for (int l = 0; l <= nd; ++l) {
        P.x_off[l] = off; off = align_up(off + 4ll * B * P.h[l] * P.w[l] * P.Ctot[l], 256);
        if (P.Ctot[l] > P.maxC) P.maxC = P.Ctot[l];
    }
    for (int l = 0; l <= nd; ++l) { P.stat_off[l] = off; off = align_up(off + 16ll * P.G * P.Ctot[l], 256); }
    for (int l = 0; l <= nd; ++l) { P.mi_off[l] = off; off = align_up(off + 8ll * P.G * P.Ctot[l], 256); }
    auto coef = [&](BnP& q) { q.coef = off / 4; off = align_up(off + 16ll * P.G * q.c, 256); };   // [G][C][4]
    for (int l = 0; l <= nd; ++l) for (auto& d : P.down[l]) coef(d.bn);
    for (int l = 0; l < nd; ++l) coef(P.td[l].bn);
    for (int i = 0; i < nd; ++i) for (auto& d : P.up[i]) coef(d.bn);
    for (int l = 0; l < nd; ++l) {
        P.td[l].argmax = off;
        off = align_up(off + 1ll * B * P.h[l + 1] * P.w[l + 1] * (P.C0[l] + P.Dn[l]), 256);
    }
    P.pre_off = off; off = align_up(off + 4ll * B * H * W, 256);
    long long maxTD = 0;                                     // widest TransitionDown (1x1 conv C -> C), rounded to 16
    for (int l = 0; l < nd; ++l) { const long long cs = (P.C0[l] + P.Dn[l] + 15) / 16 * 16; if (cs > maxTD) maxTD = cs; }
    {   // [shared slot: the transition layer being run | one image per DenseLayer, packed by ONE launch per forward]
        // 3x3 layers: one 9,216-byte image per 8-channel chunk (3xTF32); 1x1 layers: the whole matrix as hi + lo planes
        long long a = 9216ll * ((P.maxC + 7) / 8) + 9216, b2 = 8ll * maxTD * maxTD + 4096;
        long long sz = align_up(a > b2 ? a : b2, 256);
        auto slot = [&](DenseLayerP& d) { d.wp_off = sz; sz += 9216ll * ((d.cin + 7) / 8); };
        for (int l = 0; l <= nd; ++l) for (auto& d : P.down[l]) slot(d);
        for (int i = 0; i < nd; ++i) for (auto& d : P.up[i]) slot(d);
        for (int i = 0; i < nd; ++i)
            for (int q = 0; q < 8; ++q) { P.tu[i].wp_off[q] = sz; if (q * 16 < P.tu[i].conv.cout) sz += 9216ll * ((P.tu[i].cin + 7) / 8); }
        for (int q = 0; q < 8; ++q) { P.first_wp_off[q] = sz; if (q * 16 < P.first.cout) sz += 9216ll * ((c->in_channels + 7) / 8); }
        P.wpack_off = off; off = align_up(off + sz, 256);
    }
    {
        long long mx = 0;
        for (int l = 0; l < nd; ++l) {
            const long long v = 4ll * B * P.h[l] * P.w[l] * (P.C0[l] + P.Dn[l]);
            if (v > mx) mx = v;
        }
        P.tdtmp_off = off; P.tdtmp_bytes = mx; off = align_up(off + mx, 256);
    }
    P.acts_bytes = off;
    // ---- backward scratch block
    off = 0;
    for (int l = 0; l < nd; ++l) {               // in FRONT of the gradient buffers: endo_net_bwd clears from gx_off[1] to the end
        const long long n16 = 2ll * B * P.h[l] * P.w[l] * (P.C0[l] + P.Dn[l]);
        P.td[l].r16 = off; off = align_up(off + n16, 256);
        P.td[l].a16 = off; off = align_up(off + n16, 256);
    }
    {
        long long a16 = 0, g16 = 0;
        auto need = [&](const DenseLayerP& d) {
            const long long px = 1ll * B * P.h[d.level] * P.w[d.level];
            const long long a = 2 * px * ((d.cin + 7) / 8 * 8), g = 2 * px * 16;
            if (a > a16) a16 = a;
            if (g > g16) g16 = g;
        };
        for (int l = 0; l <= nd; ++l) for (auto& d : P.down[l]) need(d);
        for (int i = 0; i < nd; ++i) for (auto& d : P.up[i]) need(d);
        for (int i = 0; i < nd; ++i) {           // TransitionUp convolutions: upsampled input / all output-gradient channels at the fine level
            const long long px = 1ll * B * P.h[P.tu[i].dst_level] * P.w[P.tu[i].dst_level];
            const long long a = 2 * px * ((P.tu[i].cin + 7) / 8 * 8), g = 2 * px * ((P.tu[i].conv.cout + 15) / 16 * 16);
            if (a > a16) a16 = a;
            if (g > g16) g16 = g;
        }
        {                                        // first convolution: 8-channel image planes, all output-gradient channels
            const long long px = 1ll * B * H * W, g = 2 * px * ((P.first.cout + 15) / 16 * 16);
            if (2 * px * 8 > a16) a16 = 2 * px * 8;
            if (g > g16) g16 = g;
        }
        for (int k = 0; k < 2; ++k) {
            P.a16_off[k] = off; off = align_up(off + a16, 256);
            P.g16_off[k] = off; off = align_up(off + g16, 256);
        }
    }
    for (int l = 0; l <= nd; ++l) { P.gx_off[l] = off; off = align_up(off + 4ll * B * P.h[l] * P.w[l] * P.Ctot[l], 256); }
    for (int l = 0; l <= nd; ++l) { P.ab_off[l] = off; off = align_up(off + 8ll * P.G * P.Ctot[l], 256); }
    P.bnred_off = off; off = align_up(off + 16ll * P.G * P.maxC, 256);
    {   // [shared slot | one data-gradient image per DenseLayer]; 3x3 layers: 36,864 bytes per 64-channel chunk; 1x1: transposed matrix
        long long a = 36864ll * ((P.maxC + 63) / 64) + 36864, b2 = 4ll * maxTD * maxTD + 4096;
        long long sz = align_up(a > b2 ? a : b2, 256);
        auto slot = [&](DenseLayerP& d) { d.wpb_off = sz; sz += 36864ll * ((d.cin + 63) / 64); };
        for (int l = 0; l <= nd; ++l) for (auto& d : P.down[l]) slot(d);
        for (int i = 0; i < nd; ++i) for (auto& d : P.up[i]) slot(d);
        for (int i = 0; i < nd; ++i)
            for (int q = 0; q < 8; ++q) { P.tu[i].wpb_off[q] = sz; if (q * 16 < P.tu[i].conv.cout) sz += 36864ll * ((P.tu[i].cin + 63) / 64); }
        P.wpack_bwd_off = off; off = align_up(off + sz, 256);
    }
    P.scratch_bytes = off;
    return ENDO_OK;
}

}  // namespace endo
