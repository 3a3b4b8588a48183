// dd_layout.cuh -- caller-owned workspace layouts (shared by the sizing entry point and the kernels).
#pragma once
#include <stddef.h>

namespace dd {

// ---- synthesis -----------------------------------------------------------------------------------
constexpr int kSynthMaxBlocks = 148 * 8;  // one double partial per CTA
// (the fused synthesis + resize pass keeps one partial per CTA: at most B*3*128 for fp32 sources, one band of 2 output rows each)
inline size_t synth_ws_bytes(int B = 0) {
    const size_t fused = (size_t)(B > 0 ? B : 0) * 3 * 128;
    return sizeof(double) * (fused > (size_t)kSynthMaxBlocks ? fused : (size_t)kSynthMaxBlocks);
}
// shared memory of the fused synthesis + resize pass: the source rows behind the `rpb` output rows of a band (+1 row of the next)
constexpr size_t kSynthResizeMaxSmem = 200 * 1024;
constexpr size_t kSynthResizeTableBytes = 256 * 16 * 2 * sizeof(float);  // uint8 sources: {dark, dark - clean}, 16 bank-interleaved copies
constexpr int kResizeRowsPerBandF32 = 2, kResizeRowsPerBandU8 = 4;  // partials: at most B*3*128 (synth_ws_bytes)
inline size_t synth_resize_smem_bytes(int H, int W, int rpb) {
    const int rows = (int)((rpb + 0.5) * H / 256.0) + 3;  // band 0 also owns the rows above the first tap row (half a step)
    return (size_t)rows * W * sizeof(float);
}

// ---- dark-channel prior (train.py:42-68,81-97) -------------------------------------------------
// workspace: partial histograms [B][kPriorChunks][4][256] uint32 ({count, sum R, sum G, sum B} per dark-channel value and CTA),
// then the atmospheric light in uint8 units [B][4] float
constexpr int kPriorChunks = 32;   // CTAs per image in the histogram pass
inline size_t prior_ws_bytes(int B) { return (size_t)(B > 0 ? B : 0) * ((size_t)kPriorChunks * 4 * 256 * sizeof(unsigned) + 4 * sizeof(float)); }

// ---- predictor (common.py:52-78): channels 3->16->32->32->32->32, spatial 256->128->64->32->16->8 --
constexpr int kPredLayers = 5;
__host__ __device__ constexpr int pred_cin(int l) { return l == 0 ? 3 : l == 1 ? 16 : 32; }
__host__ __device__ constexpr int pred_cout(int l) { return l == 0 ? 16 : 32; }
__host__ __device__ constexpr int pred_hin(int l) { return 256 >> l; }
__host__ __device__ constexpr int pred_hout(int l) { return 128 >> l; }
__host__ __device__ constexpr size_t pred_act_elems(int l) {  // per image, output of conv layer l
    return (size_t)pred_cout(l) * pred_hout(l) * pred_hout(l);
}
constexpr int kFc1In = 2048, kFc1Out = 64;
// activation workspace: [a0 | a1 | a2 | a3 | a4 | h], each block is a contiguous [B, ...] tensor
inline size_t pred_act_offset(int l, int B) {  // l in 0..5 (5 = fc1 output h)
    size_t off = 0;
    for (int i = 0; i < l; ++i) off += pred_act_elems(i) * B;
    return off;
}
inline size_t predictor_acts_elems(int B) { return pred_act_offset(5, B) + (size_t)kFc1Out * B; }

// prepared weights of conv2..conv5 (dd_conv_tc.cuh: TF32 hi/lo halves in the shared-memory operand layouts of the
// forward and data-gradient GEMMs) are kept behind the activations: written by dd_predictor_fwd, read by
// dd_predictor_bwd of the same step.
__host__ __device__ constexpr size_t pred_prep_fwd_elems(int l) { return (size_t)18 * pred_cin(l) * pred_cout(l); }
__host__ __device__ constexpr size_t pred_prep_dgrad_elems(int l) { return (size_t)32 * pred_cin(l) * pred_cout(l); }
inline size_t pred_prep_offset(int l) {  // l in 1..5 (5 = end); [fwd_l | dgrad_l] per layer
    size_t off = 0;
    for (int i = 1; i < l; ++i) off += pred_prep_fwd_elems(i) + pred_prep_dgrad_elems(i);
    return off;
}
inline size_t predictor_acts_ws_elems(int B) { return predictor_acts_elems(B) + pred_prep_offset(5); }
inline size_t predictor_acts_bytes(int B) { return predictor_acts_ws_elems(B) * sizeof(float); }

// backward scratch: gradients w.r.t. the pre-activations of every layer (same layout as acts) followed by the slice
// buffers of the weight-gradient kernels (all kept until the single deferred reduction at the end of the backward).
//   conv1 (CUDA cores): 64 tiles per image x (432 + 16);  conv2..conv5 (tensor cores): one slice [RP][32] per
//   persistent CTA (at most kTcMaxCtas), RP = rows (9*Cin + 1 ones row for the bias) rounded up to 4.
constexpr int kWgradC1Slices = 64;
constexpr int kTcMaxCtas = 148;  // persistent tensor-core kernels run one CTA per SM
__host__ __device__ constexpr int pred_wgrad_rp(int l) { return (9 * pred_cin(l) + 1 + 3) / 4 * 4; }
inline size_t pred_wgrad_partial_offset(int l, int B) {  // l in 0..5 (5 = end)
    size_t off = 0;
    for (int i = 0; i < l; ++i)
        off += i == 0 ? (size_t)B * kWgradC1Slices * (432 + 16) : (size_t)kTcMaxCtas * pred_wgrad_rp(i) * 32;
    return off;
}
inline size_t predictor_bwd_ws_bytes(int B) {
    return (predictor_acts_elems(B) + pred_wgrad_partial_offset(5, B)) * sizeof(float);
}

// ---- peer exchange of the predictor gradient ----------------------------------------------
// canonical flat order = state-dict order: [w0 b0 w1 b1 w2 b2 w3 b3 w4 b4 | fc1_w fc1_b fc2_w fc2_b], 164 943 floats
constexpr int kNumGrads = 164943;
constexpr int kGradPad = 164944;  // slot pitch (floats), 16-byte multiple
__host__ __device__ constexpr int grad_off_conv_w(int l) {
    return l == 0 ? 0 : l == 1 ? 448 : 448 + 4640 + (l - 2) * 9248;
}
__host__ __device__ constexpr int grad_off_conv_b(int l) { return grad_off_conv_w(l) + 9 * pred_cin(l) * pred_cout(l); }
constexpr int kGradOffFc1W = 448 + 4640 + 3 * 9248, kGradOffFc1B = kGradOffFc1W + kFc1Out * kFc1In;
constexpr int kGradOffFc2W = kGradOffFc1B + kFc1Out, kGradOffFc2B = kGradOffFc2W + 15 * kFc1Out;
static_assert(kGradOffFc2B + 15 == kNumGrads, "flat gradient layout");
constexpr int kMaxPeers = 8;
// exchange buffer of one rank: header (256 B) + slots[2 parities][kMaxPeers][kGradPad] of 8-byte words {value, tag}
struct ExchangeHeader {
    unsigned int epoch;        // completed exchanges on this rank; the exchange in progress carries tag epoch + 1, parity epoch & 1
    unsigned int blocks_done;  // last-block counter of the exchange kernel
};
constexpr size_t kExchangeHeaderBytes = 256;
inline size_t exchange_bytes() { return kExchangeHeaderBytes + (size_t)2 * kMaxPeers * kGradPad * 8; }

// ---- fused filter chain --------------------------------------------------------------------------
// Work space of dd_recovery_fwd/bwd: every (image plane b*3+ch, 128-column strip) -- a "plane-strip" -- is cut
// into row-blocks of 32 output rows.  The N = nPS * nRB row-blocks are dealt out in contiguous, equal ranges to
// G persistent CTAs (2 per SM); a CTA marches down consecutive row-blocks of one plane-strip without re-staging
// the 24 halo rows, so load balance is +-1 row-block at any batch size and the halo is paid once per CTA.
constexpr int kStripW = 128;        // output columns per strip
constexpr int kRowBlock = 32;       // output rows per row-block
constexpr int kMaxSegRows = 640;    // longest run marched with one row-scalar table (multiple of kRowBlock)
constexpr int kSchedCtas = 148 * 2; // persistent CTAs (B200: 148 SMs x 2 resident CTAs)
constexpr int kBwdSums = 5;         // partial sums per (CTA, plane-strip): dp, dc, dgamma, ds_ch, dw

struct Sched {
    int strips;    // column strips per plane
    int nRB;       // row-blocks per plane-strip
    int nPS;       // plane-strips = B * 3 * strips
    int G;         // CTAs
    long long N;   // row-blocks in total
};

__host__ __device__ inline Sched make_sched(int B, int H, int W, int ctas = kSchedCtas) {
    Sched s;
    s.strips = (W + kStripW - 1) / kStripW;
    s.nRB = (H + kRowBlock - 1) / kRowBlock;
    s.nPS = B * 3 * s.strips;
    s.N = (long long)s.nPS * s.nRB;
    s.G = (int)(s.N < ctas ? s.N : ctas);
    return s;
}
// first row-block of CTA c (c == G gives N)
__host__ __device__ inline long long sched_begin(const Sched& s, int c) { return s.N * c / s.G; }

// backward workspace: [G + nPS][kBwdSums] floats of partial sums -- slot (c + ps) belongs to the pair (CTA c,
// plane-strip ps), which is injective because CTA ranges are ordered -- then [B*3][H][strips] floats of
// per-row partial sums S (the row-coupled contrast term).
// (the tensor-core variant keeps four row sums per (row, strip), one per 32-column quarter; the workspace covers both)
inline size_t recovery_bwd_ws_bytes(int B, int H, int W) {
    const Sched s = make_sched(B, H, W);
    return ((size_t)(s.G + s.nPS) * kBwdSums + (size_t)B * 3 * H * s.strips * 4) * sizeof(float);
}

}  // namespace dd
