// dd_layout.cuh -- caller-owned workspace layouts (shared by the sizing entry point and the kernels).
#pragma once
#include <stddef.h>

namespace dd {

// ---- synthesis -----------------------------------------------------------------------------------
constexpr int kSynthMaxBlocks = 148 * 8;  // one double partial per CTA
inline size_t synth_ws_bytes() { return sizeof(double) * kSynthMaxBlocks; }

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
inline size_t predictor_acts_bytes(int B) { return predictor_acts_elems(B) * sizeof(float); }

// backward scratch: gradients w.r.t. the pre-activations of every layer (same layout as acts) followed by
// the split-K partial buffer of the weight-gradient kernels.
constexpr int kWgradMaxSplit = 32;
constexpr size_t kWgradPartialElems = (size_t)kWgradMaxSplit * (32 * 32 * 9 + 32);
inline size_t predictor_bwd_ws_bytes(int B) {
    return (predictor_acts_elems(B) + kWgradPartialElems) * sizeof(float);
}

// ---- fused filter chain --------------------------------------------------------------------------
// Work decomposition of dd_recovery_fwd/bwd: a *unit* is (image plane b*3+ch, column strip, row segment).
constexpr int kStripW = 128;       // output columns per strip
constexpr int kMaxSegRows = 512;   // max output rows per segment (bounds the per-CTA row-scalar table)
constexpr int kBwdSums = 5;        // per-unit partial sums: dp, dc, dgamma, ds_ch, dw

struct RecoveryGrid {
    int strips;     // column strips per plane
    int segs;       // row segments per plane
    int seg_rows;   // rows per segment (last one may be shorter)
    int units;      // B * 3 * strips * segs
};

inline RecoveryGrid recovery_grid(int B, int H, int W) {
    RecoveryGrid g;
    g.strips = (W + kStripW - 1) / kStripW;
    // enough units to fill 148 SMs a few times over, but keep segments tall (each costs 24 halo rows)
    int segs = 1;
    const long long planes_strips = (long long)B * 3 * g.strips;
    while (planes_strips * segs < 148 * 4 && (H + segs - 1) / segs > 96) ++segs;
    while ((H + segs - 1) / segs > kMaxSegRows) ++segs;
    g.segs = segs;
    g.seg_rows = (H + segs - 1) / segs;
    g.segs = (H + g.seg_rows - 1) / g.seg_rows;
    g.units = (int)(planes_strips * g.segs);
    return g;
}

// backward workspace: [units][kBwdSums] floats of per-unit partial sums, then [B*3][H][strips] floats of
// per-row partial sums S (the row-coupled contrast term).
inline size_t recovery_bwd_ws_bytes(int B, int H, int W) {
    const RecoveryGrid g = recovery_grid(B, H, W);
    return ((size_t)g.units * kBwdSums + (size_t)B * 3 * H * g.strips) * sizeof(float);
}

}  // namespace dd
