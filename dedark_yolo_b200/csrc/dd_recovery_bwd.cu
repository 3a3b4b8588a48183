// dd_recovery_bwd.cu -- a14: backward of the fused filter chain.  g = dL/dy -> dfeat [B,15] (+ optional dL/dx).
//
// Closed form (checked against autograd): with x1..x4 the chain intermediates,
//   g4 = g5 (1+p) - p B^T g5        dp = sum x4 (g5 - B^T g5)        dc = sum g4 x3 (q - 1)
//   g3 = g4 m (+ row-coupled term)  dgamma = sum g3 x3 ln x2c        g2 = g3 gamma x3 / x2c [x2 >= 1e-4]
//   ds_c = sum g2 x1                g1 = g2 s_c                      dw = sum g1 (x0 - A) IcA / tx^2 [tx >= 0.01]
// where B^T is the ADJOINT of the reflect-padded 25x25 blur: a zero-padded correlation plus the reflect fold-back,
// written as a gather  W(i,j) = k[j-i] + [j>=1] k[j+i] + [j<=n-2] k[2(n-1)-j-i]  (only rows/columns within 12 of the
// image border see the extra terms).  Nothing is saved by the forward: the chain is recomputed from x.
//
// Same marching-strip skeleton as the forward (dd_recovery_fwd.cu), per 32-row block:
//   0. g block n+1 streams into the 80-row ring XS with cp.async (16 B, zero-fill outside the image) while block n
//      is processed; the x0 values of this block's output rows are prefetched into registers
//   1. H adjoint pass (FFMA2)                                                                        -> ring HS
//   2. V adjoint pass (FFMA2) -> B^T g5 for 32 rows x 128 columns                                   -> tile BT
//   3. pointwise: one warp per output row (float4 per lane): chain recompute, the five parameter sums in registers,
//      the row sum S = sum_w g4 x3 (the rgb2lum quirk couples every pixel of a row to columns 0..2) by warp shuffle.
// Reflect borders (round 2): the adjoint of the reflect-padded blur equals the PLAIN 25-tap pass over the mirror-extended cotangent
// g^ (g^[-i] = g[i], the image-border row / column doubled) with the border outputs halved (dd_blur_tc.cuh derives it; it holds
// when the two borders are more than one radius apart, H, W >= 14).  The g ring is completed in place right after a block's rows
// have landed -- mirrored halo columns, doubled border column, mirrored rows above / below the image, doubled border rows, spread
// over all 256 threads -- and both passes then run without any border code.  The previous form (zero-filled halo + the fold-back
// gather hfold8 / vfold on the outputs within 12 of a border) put several hundred serial instructions on one or two warps of every
// block of a border strip while the other warps waited at the barrier: 135 us against 98 us without it (profiles/r02_*), i.e. the
// per-SM busy time ranged from 156 k to 247 k cycles.  Images with H or W = 13 keep the fold-back form.
// Per (CTA, plane-strip) partial sums and per-row S go to the workspace; a fixed-order finalize kernel (one CTA per
// image, no float atomics) adds them up, applies the column 0..2 fix-up and the regressor Jacobians.
#include <cooperative_groups.h>
#include <cstring>
#include <type_traits>

#include "dd_async.cuh"
#include "dd_recovery.cuh"

namespace dd {

constexpr int kXRingB = 80;   // XS ring depth of the backward kernel: 32 (H pass) + 12 (lagging centre rows) + 32 in flight
constexpr int kBP = kStripW;  // BT pitch (floats)
constexpr int kFinThreads = 512;   // finalize: a cluster of kFinCluster CTAs per image, ~1 row of the column 0..2 fix-up per thread
constexpr int kFinCluster = 4;
constexpr int kPW4 = (kRB * kStripW / 4) / kThreads;  // float4 per thread in the pointwise phase (4)

__constant__ float c_tap[13] = {DD_K0, DD_K1, DD_K2, DD_K3, DD_K4, DD_K5, DD_K6,
                                DD_K7, DD_K8, DD_K9, DD_K10, DD_K11, DD_K12};

// Reflect fold-back of the adjoint blur, W(i,j) = k[j-i] + [j>=1] k[j+i] + [j<=n-2] k[2(n-1)-j-i]: the extra two terms,
// only for outputs within 12 of an image border.  Kept out of line so the hot loop stays small (instruction cache).
__device__ __noinline__ void hfold8(float* o, const float* __restrict__ xrow, int j0, int c0, int W) {
    for (int t = 0; t < 8; ++t) {
        const int j = j0 + t;
        float a = o[t];
        if (j >= 1 && j <= kRadius)
            for (int i = 0; i <= kRadius - j; ++i) a = fmaf(xrow[i - c0 + kRadius], c_tap[j + i], a);
        if (j >= W - 1 - kRadius && j <= W - 2)
            for (int i = max(0, 2 * (W - 1) - j - kRadius); i <= W - 1; ++i)
                a = fmaf(xrow[i - c0 + kRadius], c_tap[abs(2 * (W - 1) - j - i)], a);
        o[t] = a;
    }
}
__device__ __noinline__ float2 vfold(float2 bt, const float* __restrict__ HS, int col2, int jr, int r0, int H) {
    if (jr >= 1 && jr <= kRadius)
        for (int i = 0; i <= kRadius - jr; ++i) {
            const float2 h = *reinterpret_cast<const float2*>(HS + ((i - r0 + kRadius) & (kHRing - 1)) * kHP + col2);
            bt.x = fmaf(h.x, c_tap[jr + i], bt.x);
            bt.y = fmaf(h.y, c_tap[jr + i], bt.y);
        }
    if (jr >= H - 1 - kRadius && jr <= H - 2)
        for (int i = max(0, 2 * (H - 1) - jr - kRadius); i <= H - 1; ++i) {
            const float2 h = *reinterpret_cast<const float2*>(HS + ((i - r0 + kRadius) & (kHRing - 1)) * kHP + col2);
            const float kk = c_tap[abs(2 * (H - 1) - jr - i)];
            bt.x = fmaf(h.x, kk, bt.x);
            bt.y = fmaf(h.y, kk, bt.y);
        }
    return bt;
}

struct BwdAcc {
    float p, c, g, s, w;
};

// one pixel of the pointwise backward; returns dL/dx0 through the chain, adds g4*x3 to srow
template <bool HAS_ICA, bool FAST>
__device__ __forceinline__ float px_bwd(float x0, float ica, float g5, float bt, float m, float q1, const ChainK& ck,
                                        float pp, BwdAcc& acc, float& srow) {
    if (!HAS_ICA) {
        // default IcA: DeDark + WB are one FMA (the forward's own expression), and the two sums that need x1 and x0 - a are
        // taken in terms of x2:  with h2 = g2 / gamma,  acc.s = sum h2 x2,  acc.w = sum h2;  the finalize kernel turns them
        // into  ds = gamma acc.s / s  and  dw = gamma (acc.s - s a acc.w) / inv  (x1 = x2 / s, x0 - a = (x1 - a) / inv):
        // 4 instructions per pixel less than carrying x1 and x0 - a through the chain.
        const float x2 = fmaf(x0, ck.k1, ck.k0);
        const float x2c = fmaxf(x2, kGammaClamp);
        float l2;
        const float x3 = gamma_pow<FAST>(x2c, ck.gamma, &l2);
        const float g4 = fmaf(g5, 1.f + pp, -pp * bt);
        acc.p = fmaf(x3 * m, g5 - bt, acc.p);
        const float u3 = g4 * x3;
        srow += u3;
        acc.c = fmaf(u3, q1, acc.c);
        const float g3x3 = u3 * m;
        acc.g = fmaf(g3x3, l2, acc.g);  // log2: * ln 2 applied in finalize
        const float h2 = x2 >= kGammaClamp ? g3x3 * rcp_fast(x2c) : 0.f;
        acc.s = fmaf(h2, x2, acc.s);
        acc.w += h2;
        return h2 * (ck.gamma * ck.k1);  // dL/dx0 = g2 s inv = h2 gamma k1
    }
    const float xa = x0 - ck.a;
    const float tx = fmaf(-ck.w, ica, 1.f);
    const bool pass_tx = tx >= kTxMin;
    const float txc = fmaxf(tx, kTxMin);
    const float inv = __fdiv_rn(1.f, txc);
    const float x1 = __fdiv_rn(xa, txc) + ck.a;
    const float icaw = ica * inv * inv;
    const float x2 = x1 * ck.s;
    const float x2c = fmaxf(x2, kGammaClamp);
    float l2;
    const float x3 = gamma_pow<FAST>(x2c, ck.gamma, &l2);
    const float g4 = fmaf(g5, 1.f + pp, -pp * bt);
    acc.p = fmaf(x3 * m, g5 - bt, acc.p);
    const float u3 = g4 * x3;  // g4 * x3
    srow += u3;
    acc.c = fmaf(u3, q1, acc.c);
    const float g3x3 = u3 * m;  // g3 * x3
    acc.g = fmaf(g3x3, l2, acc.g);  // log2: * ln 2 applied in finalize
    const float g2 = x2 >= kGammaClamp ? g3x3 * ck.gamma * rcp_fast(x2c) : 0.f;
    acc.s = fmaf(g2, x1, acc.s);
    const float g1 = g2 * ck.s;
    // d x1 / d w = (x0 - a) * ica / txc^2   (tx = 1 - w ica, only where tx >= 0.01)
    if (pass_tx) acc.w = fmaf(g1 * xa, icaw, acc.w);
    return g1 * inv;
}

}  // namespace dd
#include "dd_recovery_tc_bwd.cuh"  // the tensor-core variant (uses px_bwd / BwdAcc above)
namespace dd {

// TMA: the g halo tile of every row-block arrives as two 16-row x 156-column boxes (cp.async.bulk.tensor through a tensor
// map of the [planes][H][W] cotangent; everything outside the image is zero-filled by the copy engine), issued by one
// thread and counted on an mbarrier: no per-thread address arithmetic, no staging instructions at all.
// XT: element type of x -- 0 fp32; 1 uint8: the source batch, read through the 256-entry darkening table
// `dark_tab` (one 32-bit load per four pixels); 2 bf16 (the bf16 I/O mode).  GB: the cotangent g is bf16 -- it
// is converted to fp32 on its way into the ring (register prefetch one block ahead; TMA moves bytes, it cannot widen them).
// The arithmetic is fp32 throughout.  XT != 0 and GB need ALIGNED; no dx for a uint8 source.
constexpr int kXF32 = 0, kXU8 = 1, kXBF16 = 2;
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 q) {
    return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u), __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xFFFF0000u));
}
template <bool HAS_ICA, bool FAST, bool ALIGNED, bool TMA, int XT = kXF32, bool GB = false>
__global__ void __launch_bounds__(kThreads, 2)
recovery_bwd_kernel(const __grid_constant__ CUtensorMap gmap, const float* __restrict__ x, const float* __restrict__ A,
                    const float* __restrict__ IcA, const float* __restrict__ feat, const float* __restrict__ g,
                    float* __restrict__ part, float* __restrict__ Spart, float* __restrict__ dx, int B, int H, int W,
                    const float* __restrict__ dark_tab = nullptr) {
    constexpr bool U8 = XT == kXU8, XB = XT == kXBF16;
    static_assert((XT == kXF32 && !GB) || ALIGNED, "uint8 / bf16 operands need aligned rows");
    static_assert(!(GB && TMA), "a bf16 cotangent is staged through registers");
    __shared__ float s_tab[U8 ? 256 : 1];
    const unsigned char* x8 = reinterpret_cast<const unsigned char*>(x);
    const unsigned short* xb = reinterpret_cast<const unsigned short*>(x);
    const unsigned short* gb = reinterpret_cast<const unsigned short*>(g);
    auto lut4 = [&](unsigned w) { return make_float4(s_tab[w & 255u], s_tab[(w >> 8) & 255u], s_tab[(w >> 16) & 255u], s_tab[w >> 24]); };
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t tma_bar;
    uint32_t tma_phase = 0;
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&tma_bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    // As in the forward kernel: the batch and the cotangent (x, IcA, A, g) are inputs of the step -- no kernel of the library that
    // writes image-sized tensors releases its dependents early -- so the first segment requests its per-row columns and its first
    // cotangent tile BEFORE the wait, which sits in front of the regressors (feat is the only operand a predecessor may still be
    // writing).  uint8 sources wait here (the table).
    // (wait, THEN launch_dependents: the finalize kernel behind this one reads x and feat ahead of its own wait)
    bool waited = false;
    if (U8) {
        pdl_wait();
        pdl_launch();
        waited = true;
        if (threadIdx.x < 256) s_tab[threadIdx.x] = __ldcg(dark_tab + threadIdx.x);   // not __ldg: ptxas hoists ld.global.nc above the wait
        __syncthreads();
    }
    float* XS = smem;                   // g, zero outside the image
    float* HS = XS + kXRingB * kXP;
    float* BT = HS + kHRing * kHP;      // B^T g5 for the 32 output rows of the current block
    float* MSm = BT + kRB * kBP;        // per virtual row: m = (1-c) + c*q
    float* MSq = MSm + kMaxU;           // per virtual row: q - 1
    __shared__ ImgParams sp;
    __shared__ float s_red[(kThreads / 32) * kBwdSums];

    const int tid = threadIdx.x, lane = tid & 31;
    const Sched sc = make_sched(B, H, W);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);

    BwdAcc acc = {0.f, 0.f, 0.f, 0.f, 0.f};
    int cur_ps = -1;
    auto flush = [&]() {  // per (CTA, plane-strip) partial sums -> slot (cta + ps); the five sums share two barriers
        float* out = part + (size_t)(blockIdx.x + cur_ps) * kBwdSums;
        const float v[kBwdSums] = {warp_sum(acc.p), warp_sum(acc.c), warp_sum(acc.g), warp_sum(acc.s), warp_sum(acc.w)};
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < kBwdSums; ++j) s_red[(tid >> 5) * kBwdSums + j] = v[j];
        }
        __syncthreads();
        if (tid < kBwdSums) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += s_red[w * kBwdSums + tid];  // fixed order
            out[tid] = s;
        }
        __syncthreads();
        acc.p = acc.c = acc.g = acc.s = acc.w = 0.f;
    };

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        if (cur_ps >= 0 && u.ps != cur_ps) flush();
        cur_ps = u.ps;
        __syncthreads();
        const float* xp = x + (size_t)u.plane * H * W;
        const unsigned char* xp8 = x8 + (size_t)u.plane * H * W;
        const unsigned short* xpb = xb + (size_t)u.plane * H * W;
        const unsigned short* gpb = gb + (size_t)u.plane * H * W;
        const float* gp = g + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
        // columns 0..2 of every row of the segment (the per-row contrast scalars below): requested before the regressors are
        // evaluated -- they do not depend on them, and one warp's tanhf / expf calls then run in the shadow of these loads
        constexpr int kPer = (kMaxU + kThreads - 1) / kThreads;
        float x0r[kPer][3], icr[kPer][3];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int v = tid + k * kThreads, row = u.r0 - kRadius + v;
            if (v < u.nU && row >= 0 && row < H) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    x0r[k][c] = U8 ? s_tab[__ldg(xp8 + (size_t)row * W + c)]
                                   : XB ? __uint_as_float((unsigned)__ldg(xpb + (size_t)row * W + c) << 16) : __ldg(xp + (size_t)row * W + c);
                    icr[k][c] = HAS_ICA ? __ldg(ip + (size_t)row * W + c) : kDefaultIcA;
                }
            }
        }
        // g block n -> ring XS (rows outside the image and columns outside [0, W) are zero)
        uint2 gpre[GB ? kStage4 : 1];   // GB: the bf16 rows of the next block, in flight between stage(n) and stage_commit(n)
        // GB: a thread's items of a block are the same (row, column chunk) pairs in every block -- their offsets are computed once per
        // segment, a block then costs one ring-row wrap test and one bounds test per item
        int g_rr[GB ? kStage4 : 1], g_soff[GB ? kStage4 : 1];
        unsigned g_goff[GB ? kStage4 : 1];
        bool g_colok[GB ? kStage4 : 1];
        if (GB) {
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4, gc = u.c0 - kRadius + 4 * c4;
                g_rr[k] = f < kRB * kXW4 ? rr : -1;
                g_soff[k] = 4 * c4;
                g_colok[k] = gc >= 0 && gc < W;
                g_goff[k] = (unsigned)(rr * W + gc);   // + (first row of the block) * W; only used when the row and column are inside
            }
        }
        auto stage_commit = [&](int n) {  // GB only: widen and store block n into the ring (its rows are free since the last barrier)
            if (!GB) return;
            const int rb = (n * kRB) % kXRingB;
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                if (g_rr[GB ? k : 0] >= 0) {
                    int ring = rb + g_rr[GB ? k : 0];
                    ring = ring >= kXRingB ? ring - kXRingB : ring;
                    *reinterpret_cast<float4*>(XS + ring * kXP + g_soff[GB ? k : 0]) = bf16x4_to_float4(gpre[GB ? k : 0]);
                }
            }
        };
        auto stage = [&](int n) {
            if (GB) {
                const int row0 = u.r0 - kRadius + n * kRB;
                const unsigned short* base = gpb + (ptrdiff_t)row0 * W;
#pragma unroll
                for (int k = 0; k < kStage4; ++k) {
                    const int row = row0 + g_rr[GB ? k : 0];
                    const bool ok = g_rr[GB ? k : 0] >= 0 && g_colok[GB ? k : 0] && row >= 0 && row < H;
                    gpre[GB ? k : 0] = ok ? __ldg(reinterpret_cast<const uint2*>(base + g_goff[GB ? k : 0])) : make_uint2(0u, 0u);
                }
                return;
            }
            if (TMA) {
                if (tid == 0) {
                    mbar_arrive_expect_tx(&tma_bar, 2u * 16u * kXP * 4u);
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int v0 = n * kRB + 16 * hh;
                        tma_load_3d(XS + (v0 % kXRingB) * kXP, &gmap, u.c0 - kRadius, u.r0 - kRadius + v0, u.plane, &tma_bar);
                    }
                }
                return;
            }
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4;
                if (f < kRB * kXW4) {
                    const int v = n * kRB + rr;
                    const int row = u.r0 - kRadius + v;
                    const int gc = u.c0 - kRadius + 4 * c4;
                    float* dst = XS + (v % kXRingB) * kXP + 4 * c4;
                    const bool rok = row >= 0 && row < H;
                    if (ALIGNED) {
                        const bool ok = rok && gc >= 0 && gc < W;
                        cp_async16(dst, ok ? gp + (size_t)row * W + gc : gp, ok);
                    } else {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rok) {
                            const float* rp = gp + (size_t)row * W;
                            if (gc >= 0 && gc < W) t.x = __ldg(rp + gc);
                            if (gc + 1 >= 0 && gc + 1 < W) t.y = __ldg(rp + gc + 1);
                            if (gc + 2 >= 0 && gc + 2 < W) t.z = __ldg(rp + gc + 2);
                            if (gc + 3 >= 0 && gc + 3 < W) t.w = __ldg(rp + gc + 3);
                        }
                        *reinterpret_cast<float4*>(dst) = t;
                    }
                }
            }
            cp_async_commit();
        };
        stage(0);
        if (!waited) {
            pdl_wait();
            pdl_launch();
            waited = true;
        }
        if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
        __syncthreads();
        const ChainK ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
        const float pc = sp.c, pp = sp.p;

        // ---- mirror extension of the cotangent ring (see the header) -------------------------------------------------------------
        const bool mirror = H > kRadius + 1 && W > kRadius + 1;
        const bool col_fix = mirror && (u.c0 == 0 || u.c0 + kStripW + kRadius > W);   // the tile holds columns < 0 or >= W
        const int tc_last = W - 1 - u.c0 + kRadius;                                    // tile column of image column W - 1
        const int wrow = tid >> 5, lane4 = 4 * lane;                                   // pointwise phase: this thread's row offset / column chunk
        const int gc = u.c0 + lane4;                                                   // ... and its first image column
        const ptrdiff_t w8 = 8 * (ptrdiff_t)W;                                         // eight image rows (elements)
        float fxc[4];                                                                  // pointwise phase: this thread's four columns
        bool edge_col = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gcol = u.c0 + 4 * (tid & 31) + i;
            fxc[i] = (gcol == 0 || gcol == W - 1) ? 0.5f : 1.f;
            edge_col = edge_col || gcol == 0 || gcol == W - 1;
        }
        auto mirror_fixup = [&](int n) {   // block n's rows are in the ring and visible to this thread's CTA
            const int row_first = u.r0 - kRadius + n * kRB;                           // image row of the block's first ring row
            const bool row_fix = row_first < 0 || row_first + kRB > H - 1;            // rows outside the image, or a border row
            if (!col_fix && !row_fix) return;                                         // (CTA-uniform)
            if (!TMA) __syncthreads();                                                // rows staged by other threads
            if (col_fix) {   // in-image rows: 8 threads per row, 26 independent items (12 + 12 mirrored columns, 2 doubled ones)
                const int rr = tid >> 3, jr = row_first + rr;
                if (jr >= 0 && jr < H) {
                    float* xr = XS + ((n * kRB + rr) % kXRingB) * kXP;
#pragma unroll
                    for (int k0 = 0; k0 < 32; k0 += 8) {
                        const int k = k0 + (tid & 7);
                        if (k < kRadius) {                                            // column -(k+1) <- column k+1   (strip 0 only)
                            if (u.c0 == 0) xr[kRadius - 1 - k] = xr[kRadius + 1 + k];
                        } else if (k < 2 * kRadius) {                                 // column W-1+i <- column W-1-i
                            const int i = k - kRadius + 1, td = tc_last + i, ts = tc_last - i;
                            if (td >= 0 && td < kXW && ts >= 0) xr[td] = xr[ts];
                        } else if (k == 2 * kRadius) {                                // g^[0] = 2 g[0]
                            if (u.c0 == 0) xr[kRadius] *= 2.f;
                        } else if (k == 2 * kRadius + 1) {                            // g^[W-1] = 2 g[W-1]
                            if (tc_last >= 0 && tc_last < kXW) xr[tc_last] *= 2.f;
                        }
                    }
                }
            }
            if (row_fix) {   // rows above / below the image <- their mirror rows (already column-complete); border rows doubled
                __syncthreads();
                for (int it = tid; it < kRB * kXW4; it += kThreads) {
                    const int rr = it / kXW4, c4 = it - rr * kXW4, jr = row_first + rr;
                    float4* dst = reinterpret_cast<float4*>(XS + ((n * kRB + rr) % kXRingB) * kXP) + c4;
                    if (jr < 0 || jr >= H) {
                        const int js = jr < 0 ? -jr : 2 * (H - 1) - jr;               // reflect
                        if (js >= 0 && js < H && (jr >= -kRadius && jr < H + kRadius)) {
                            const int vs = js - (u.r0 - kRadius);                     // ring row of the source (an earlier or the same block)
                            *dst = *(reinterpret_cast<const float4*>(XS + (vs % kXRingB) * kXP) + c4);
                        }
                    } else if (jr == 0 || jr == H - 1) {
                        float4 v = *dst;
                        v.x *= 2.f; v.y *= 2.f; v.z *= 2.f; v.w *= 2.f;
                        *dst = v;
                    }
                }
            }
        };

        {   // per-row contrast scalars of the segment (the rows' first three columns were loaded at the top of the segment)
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int v = tid + k * kThreads, row = u.r0 - kRadius + v;
                if (v < u.nU) {
                    float m = 0.f, q1 = 0.f;
                    if (row >= 0 && row < H) {
                        float x3[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) x3[c] = chain_x3<HAS_ICA, FAST>(ck, x0r[k][c], icr[k][c]);
                        const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
                        q1 = rl.q - 1.f;
                        m = (1.f - pc) + pc * rl.q;
                    }
                    MSm[v] = m;
                    MSq[v] = q1;
                }
            }
        }

        // Two barriers per row-block: the pointwise phase of block n-1 shares an interval with the horizontal pass of block n (both
        // only read the g ring; the pointwise phase reads BT, the horizontal pass writes the H ring), the vertical pass of block n
        // (reads the H ring, writes BT) has the other.  One extra iteration drains the last pointwise phase.
        float4 x0p[kPW4];
        for (int n = 0; n <= u.nB; ++n) {
            const bool have = n < u.nB;
            if (have) {
                if (GB) {
                    stage_commit(n);
                } else if (TMA) {
                    mbar_wait(&tma_bar, tma_phase);
                    tma_phase ^= 1u;
                } else {
                    cp_async_wait_all();
                }
                if (mirror) mirror_fixup(n);
            }
            __syncthreads();  // g block n visible; BT of block n-1 complete; the H ring rows block n overwrites are no longer read
            if (n > 0) {
                // pointwise phase: warp <-> output row, lane <-> float4.  A thread's four items of a block are rows ob, ob + 8, ob + 16,
                // ob + 24 of one column chunk: every index is a per-block base plus a constant step (no per-item multiplications)
                const int ob = (n - 1) * kRB - kRadius + wrow;             // virtual row of item 0
                int ring = ob % kXRingB;                                   // ring row of item 0 (rows with ob < 0 are skipped below)
                ring = ring < 0 ? ring + kXRingB : ring;
                const ptrdiff_t row_off0 = (ptrdiff_t)(u.r0 + ob - kRadius) * W + gc;   // image offset of item 0 (element units)
                float* sp_ptr = Spart + ((ptrdiff_t)u.plane * H + (u.r0 + ob - kRadius)) * sc.strips + u.strip;
    #pragma unroll
                for (int k = 0; k < kPW4; ++k) {
                    const int rr = wrow + 8 * k;
                    const int o = ob + 8 * k;
                    const bool row_ok = (unsigned)(o - kRadius) < (unsigned)u.seg_len;  // warp-uniform
                    int rk = ring + 8 * k;
                    rk = rk >= kXRingB ? rk - kXRingB : rk;
                    if (!row_ok) continue;
                    const int jr = u.r0 + o - kRadius;
                    float srow = 0.f;
                    if (gc < W) {
                        float4 g5 = *reinterpret_cast<const float4*>(XS + rk * kXP + lane4 + kRadius);
                        float4 bt = *reinterpret_cast<const float4*>(BT + rr * kBP + lane4);
                        const float m = MSm[o], q1 = MSq[o];
                        const float4 x0 = U8 ? lut4(__float_as_uint(x0p[k].x))
                                          : XB ? bf16x4_to_float4(make_uint2(__float_as_uint(x0p[k].x), __float_as_uint(x0p[k].y))) : x0p[k];
                        const ptrdiff_t off = row_off0 + (ptrdiff_t)k * w8;
                        float4 ic = make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                        if (HAS_ICA) {
                            ic.x = __ldg(ip + off);
                            if (gc + 1 < W) ic.y = __ldg(ip + off + 1);
                            if (gc + 2 < W) ic.z = __ldg(ip + off + 2);
                            if (gc + 3 < W) ic.w = __ldg(ip + off + 3);
                        }
                        // columns beyond W - 1 (only when W % 4 != 0): g5 and bt are zero-padded there, but bt is not
                        // (it is a blur of real data), so mask their cotangents explicitly
                        if (mirror && (edge_col || jr == 0 || jr == H - 1)) {   // border row / column: g^ was doubled there, B^T g^ is twice B^T g
                            const float fr = (jr == 0 || jr == H - 1) ? 0.5f : 1.f;
                            g5.x *= fr * fxc[0]; g5.y *= fr * fxc[1]; g5.z *= fr * fxc[2]; g5.w *= fr * fxc[3];
                            bt.x *= fr * fxc[0]; bt.y *= fr * fxc[1]; bt.z *= fr * fxc[2]; bt.w *= fr * fxc[3];
                        }
                        float4 d;
                        d.x = px_bwd<HAS_ICA, FAST>(x0.x, ic.x, g5.x, bt.x, m, q1, ck, pp, acc, srow);
                        if (ALIGNED) {  // W % 4 == 0: a float4 that starts inside the image ends inside it
                            d.y = px_bwd<HAS_ICA, FAST>(x0.y, ic.y, g5.y, bt.y, m, q1, ck, pp, acc, srow);
                            d.z = px_bwd<HAS_ICA, FAST>(x0.z, ic.z, g5.z, bt.z, m, q1, ck, pp, acc, srow);
                            d.w = px_bwd<HAS_ICA, FAST>(x0.w, ic.w, g5.w, bt.w, m, q1, ck, pp, acc, srow);
                        } else {
                            const float l1 = gc + 1 < W ? 1.f : 0.f, l2 = gc + 2 < W ? 1.f : 0.f, l3 = gc + 3 < W ? 1.f : 0.f;
                            d.y = px_bwd<HAS_ICA, FAST>(x0.y, ic.y, g5.y * l1, bt.y * l1, m, q1, ck, pp, acc, srow);
                            d.z = px_bwd<HAS_ICA, FAST>(x0.z, ic.z, g5.z * l2, bt.z * l2, m, q1, ck, pp, acc, srow);
                            d.w = px_bwd<HAS_ICA, FAST>(x0.w, ic.w, g5.w * l3, bt.w * l3, m, q1, ck, pp, acc, srow);
                        }
                        if (dx) {
                            float* dp = dx + (size_t)u.plane * H * W + off;
                            if (ALIGNED) {
                                *reinterpret_cast<float4*>(dp) = d;
                            } else {
                                dp[0] = d.x;
                                if (gc + 1 < W) dp[1] = d.y;
                                if (gc + 2 < W) dp[2] = d.z;
                                if (gc + 3 < W) dp[3] = d.w;
                            }
                        }
                    }
                    srow = warp_sum(srow);
                    if (lane == 0) sp_ptr[(ptrdiff_t)k * 8 * sc.strips] = srow;
                }
            }
            if (have) {
                {   // horizontal adjoint pass
                    const int rr = tid & 31, cg = tid >> 5;
                    const int v = n * kRB + rr;
                    const float* xrow = XS + (v % kXRingB) * kXP;
    #pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const int cb = 64 * half + 8 * cg;
                        float o[8];
                        hpass8(xrow + cb, o);
                        const int j0 = u.c0 + cb;  // global column of o[0]
                        if (!mirror && (j0 <= kRadius || j0 + 7 >= W - 1 - kRadius)) {  // reflect fold-back (H or W = 13 only)
                            float tmp[8];
    #pragma unroll
                            for (int t = 0; t < 8; ++t) tmp[t] = o[t];
                            hfold8(tmp, xrow, j0, u.c0, W);
    #pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] = tmp[t];
                        }
                        float4* dst = reinterpret_cast<float4*>(HS + (v & (kHRing - 1)) * kHP + cb);
                        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
            __syncthreads();  // H ring rows of block n visible; BT, the g centre rows and x0p of block n-1 consumed
            if (have) {
                if (n + 1 < u.nB) stage(n + 1);
                // prefetch x0 for this block's output rows (issued after the H pass to keep its register footprint down; consumed after the
                // V pass): one 64-bit base per block, a constant step of 8 rows between the four items
                {
                    const int ob = n * kRB - kRadius + wrow;
                    const ptrdiff_t off0 = (ptrdiff_t)(u.r0 + ob - kRadius) * W + gc;
    #pragma unroll
                    for (int k = 0; k < kPW4; ++k) {
                        x0p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if ((unsigned)(ob + 8 * k - kRadius) < (unsigned)u.seg_len && gc < W) {
                            const ptrdiff_t off = off0 + (ptrdiff_t)k * w8;
                            if (U8) {
                                x0p[k].x = __uint_as_float(__ldg(reinterpret_cast<const unsigned*>(xp8 + off)));
                            } else if (XB) {
                                const uint2 q = __ldg(reinterpret_cast<const uint2*>(xpb + off));
                                x0p[k].x = __uint_as_float(q.x);
                                x0p[k].y = __uint_as_float(q.y);
                            } else if (ALIGNED) {
                                x0p[k] = __ldg(reinterpret_cast<const float4*>(xp + off));
                            } else {
                                const float* rp = xp + off;
                                x0p[k].x = __ldg(rp);
                                if (gc + 1 < W) x0p[k].y = __ldg(rp + 1);
                                if (gc + 2 < W) x0p[k].z = __ldg(rp + 2);
                                if (gc + 3 < W) x0p[k].w = __ldg(rp + 3);
                            }
                        }
                    }
                }

                {   // vertical adjoint pass -> BT
                    const int col2 = 2 * (tid & 63), rg = tid >> 6;
                    const int o_first = n * kRB - kRadius + 8 * rg;
                    if (o_first >= kRadius && o_first < kRadius + u.seg_len) {
                        u64 bt2[8];
                        vpass8x2(HS, (o_first - kRadius) & (kHRing - 1), col2, bt2);
    #pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const int o = o_first + r;
                            const int jr = u.r0 + o - kRadius;  // image row
                            if (!mirror && (jr <= kRadius || jr >= H - 1 - kRadius)) {  // reflect fold-back (H or W = 13 only)
                                const float2 bt = vfold(upk(bt2[r]), HS, col2, jr, u.r0, H);
                                bt2[r] = pk(bt.x, bt.y);
                            }
                            *reinterpret_cast<u64*>(BT + (8 * rg + r) * kBP + col2) = bt2[r];
                        }
                    }
                }
            }
        }
    }
    if (cur_ps >= 0) flush();
}

// x0 of the column 0..2 fix-up: fp32 / bf16 images are read directly, a uint8 source through the darkening table
template <typename TX>
__device__ __forceinline__ float load_x0(const TX* x, size_t i, const float*) { return btc::Elem<TX>::load1(x + i); }
template <>
__device__ __forceinline__ float load_x0<unsigned char>(const unsigned char* x, size_t i, const float* tab) { return __ldg(tab + __ldg(x + i)); }
// dx was written by the main kernel (the grid dependency): read it with ld.global.cg, which is never moved across the wait
template <typename TX>
__device__ __forceinline__ void add_dx(TX* dx, size_t i, float v);
template <>
__device__ __forceinline__ void add_dx<float>(float* dx, size_t i, float v) { dx[i] = __ldcg(dx + i) + v; }
template <>
__device__ __forceinline__ void add_dx<__nv_bfloat16>(__nv_bfloat16* dx, size_t i, float v) {
    const float old = __uint_as_float((unsigned)__ldcg(reinterpret_cast<const unsigned short*>(dx + i)) << 16);
    dx[i] = __float2bfloat16_rn(old + v);
}
template <>
__device__ __forceinline__ void add_dx<unsigned char>(unsigned char*, size_t, float) {}   // a uint8 source has no gradient

// One CLUSTER of 4 CTAs per image (the kernel is a chain of latencies on ~2000 rows, one CTA per image left 132 SMs idle):
// fixed-order sum of the (CTA, plane-strip) partials, the row-coupled fix-up of columns 0..2 (d lum / d x3[:, :, :, 0..2]),
// each CTA on a quarter of the rows; the four partial results meet through distributed shared memory in rank order and
// CTA 0 applies the regressor Jacobians -> dfeat[b, 0..14].
template <bool HAS_ICA, bool FAST, typename TX = float>
__global__ void __launch_bounds__(kFinThreads)
recovery_bwd_finalize_kernel(const TX* __restrict__ x, const float* __restrict__ A,
                             const float* __restrict__ IcA, const float* __restrict__ feat,
                             const float* __restrict__ part, const float* __restrict__ Spart,
                             float* __restrict__ dfeat, TX* __restrict__ dx, int B, int H, int W, int ctas, int nsp,
                             const float* __restrict__ dark_tab = nullptr) {
    // The main kernel in front releases this one only after its own wait has returned, so x, IcA, A and feat are final when a CTA
    // of this kernel starts: the regressors and the rows' first three columns are requested AHEAD of the grid dependency, the
    // main kernel's partial sums (part, Spart) behind it, with ld.global.cg.  (uint8 sources wait first: the table.)
    pdl_launch();
    constexpr bool kEarly = !std::is_same<TX, unsigned char>::value;
    if (!kEarly) pdl_wait();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ ImgParams sp;
    __shared__ double s_red[kFinThreads / 32][7];
    const int tid = threadIdx.x, b = blockIdx.x / kFinCluster, crank = (int)cluster.block_rank();
    const Sched sc = make_sched(B, H, W, ctas);  // the work list of the kernel that produced the partials (ctas CTAs, nsp row sums per row)
    const int rows_per = (3 * H + kFinCluster - 1) / kFinCluster, row_lo = crank * rows_per, row_hi = min(3 * H, row_lo + rows_per);
    if (tid < 32) regress_warp(feat + b * kFeat, sp);
    double dp = 0, dc = 0, dg = 0, dw = 0, ds[3] = {0, 0, 0};
    const float kappa[3] = {kLumR, kLumG, kLumB};

    // rows of the column 0..2 fix-up, two per thread and pass: every load is issued before anything that needs the
    // regressed parameters (this kernel is a chain of dependent latencies, not of arithmetic)
    for (int i0 = row_lo; i0 < max(row_hi, row_lo + 1); i0 += 2 * kFinThreads) {
        float S[2], x0[2][3], ica[2][3];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = i0 + e * kFinThreads + tid;
            if (i < row_hi) {
                const int ch = i / H, row = i - ch * H, plane = 3 * b + ch;
                const size_t off = ((size_t)plane * H + row) * W;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    x0[e][k] = load_x0(x, off + k, dark_tab);
                    ica[e][k] = HAS_ICA ? __ldg(IcA + ((size_t)b * H + row) * W + k) : kDefaultIcA;
                }
            }
        }
        if (kEarly && i0 == row_lo) pdl_wait();
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = i0 + e * kFinThreads + tid;
            S[e] = 0.f;
            if (i < row_hi) {
                const int ch = i / H, row = i - ch * H, plane = 3 * b + ch;
                for (int st = 0; st < nsp; ++st) S[e] += __ldcg(Spart + ((size_t)plane * H + row) * nsp + st);
            }
        }
        if (i0 == row_lo) __syncthreads();  // sp
        const float pg = sp.gamma, pc = sp.c;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = i0 + e * kFinThreads + tid;
            if (i >= row_hi) continue;
            const int ch = i / H, row = i - ch * H, plane = 3 * b + ch;
            const float a = A ? __ldg(A + b * 3 + ch) : kDefaultA;
            const ChainK ck = make_chain(sp, ch, a);
            const size_t off = ((size_t)plane * H + row) * W;
            float tx[3], txc[3], x1[3], x2[3], x2c[3], x3[3], l2[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tx[k] = 1.f - ck.w * ica[e][k];
                txc[k] = fmaxf(tx[k], kTxMin);
                x1[k] = (x0[e][k] - a) / txc[k] + a;
                x2[k] = x1[k] * ck.s;
                x2c[k] = fmaxf(x2[k], kGammaClamp);
                x3[k] = gamma_pow<FAST>(x2c[k], pg, &l2[k]);
            }
            const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
            if (rl.lraw >= 0.f && rl.lraw <= 1.f) {
                const float dq = 0.5f * kPi * sinf(kPi * rl.lum) / rl.denom - rl.cl / (rl.denom * rl.denom);
                const float glum = pc * dq * S[e];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float e3 = kappa[k] * glum;
                    dg += (double)(e3 * x3[k] * l2[k]) * 0.69314718055994530942;
                    const float e2 = x2[k] >= kGammaClamp ? e3 * pg * x3[k] / x2c[k] : 0.f;
                    ds[ch] += (double)(e2 * x1[k]);
                    const float e1 = e2 * ck.s;
                    if (tx[k] >= kTxMin) dw += (double)(e1 * (x0[e][k] - a) * ica[e][k] / (txc[k] * txc[k]));
                    if (dx) add_dx(dx, off + k, e1 / txc[k]);
                }
            }
        }
    }

    // partial sums: plane-strip ps of this image was processed by CTAs c_of(first block) .. c_of(last block)
    for (int i = crank * kFinThreads + tid; i < 3 * sc.strips; i += kFinCluster * kFinThreads) {
        const int ps = 3 * b * sc.strips + i;
        const int ch = i / sc.strips;
        const long long x0b = (long long)ps * sc.nRB, x1b = x0b + sc.nRB - 1;
        const int c_first = (int)(((x0b + 1) * sc.G + sc.N - 1) / sc.N) - 1;
        const int c_last = (int)(((x1b + 1) * sc.G + sc.N - 1) / sc.N) - 1;
        // default-IcA constants folded out of the kernel's acc.w
        const float txc = fmaxf(1.f - sp.w * kDefaultIcA, kTxMin);
        const float wk = HAS_ICA ? 1.f : ((1.f - sp.w * kDefaultIcA >= kTxMin) ? kDefaultIcA / (txc * txc) : 0.f);
        const ChainK ckp = make_chain(sp, ch, A ? __ldg(A + b * 3 + ch) : kDefaultA);
        for (int c = c_first; c <= c_last; ++c) {
            const float* qp = part + (size_t)(c + ps) * kBwdSums;
            const float q[kBwdSums] = {__ldcg(qp), __ldcg(qp + 1), __ldcg(qp + 2), __ldcg(qp + 3), __ldcg(qp + 4)};
            dp += q[0]; dc += q[1]; dg += (double)q[2] * 0.69314718055994530942;
            if (HAS_ICA) {
                ds[ch] += q[3];
                dw += (double)q[4];
            } else {  // q[3] = sum h2 x2, q[4] = sum h2 (see px_bwd)
                ds[ch] += (double)ckp.gamma * (double)q[3] / (double)ckp.s;
                dw += (double)ckp.gamma * ((double)q[3] - (double)ckp.s * (double)ckp.a * (double)q[4]) / (double)ckp.inv * (double)wk;
            }
        }
    }
    {   // the seven block sums in one pass: lanes (shuffles), then warps in index order (deterministic)
        double v[7] = {dp, dc, dg, dw, ds[0], ds[1], ds[2]};
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            v[k] = warp_sum(v[k]);
            if (lane == 0) s_red[wid][k] = v[k];
        }
        __syncthreads();
        if (tid < 7) {
            double r = 0.0;
            for (int w = 0; w < kFinThreads / 32; ++w) r += s_red[w][tid];
            s_red[0][tid] = r;
        }
        __syncthreads();
    }
    cluster.sync();
    if (crank == 0 && tid == 0) {  // the four quarters in rank order (deterministic)
        double v[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int r = 0; r < kFinCluster; ++r) {
            const double* q = cluster.map_shared_rank(&s_red[0][0], r);
            for (int k = 0; k < 7; ++k) v[k] += q[k];
        }
        dp = v[0]; dc = v[1]; dg = v[2]; dw = v[3];
        ds[0] = v[4]; ds[1] = v[5]; ds[2] = v[6];
    }
    if (crank == 0 && tid == 0) {
        float* o = dfeat + b * kFeat;
        for (int i = 0; i < kFeat; ++i) o[i] = 0.f;
        const float* t = sp.t;
        o[kSlotDedark] = (float)(dw * 0.45 * (1.0 - (double)t[0] * t[0]));
        // WB: s_c = cs_c / Z, cs_j = exp(0.5 t_j), t_j = tanh(f_{1+j} m_j), m = (0,1,1)
        const double Z = sp.Z;
        double dot = 0.0;
        for (int c2 = 0; c2 < 3; ++c2) dot += ds[c2] * (double)sp.cs[c2];
        for (int j = 1; j < 3; ++j) {
            const double dcs = ds[j] / Z - (double)kappa[j] * dot / (Z * Z);
            o[kSlotWb + j] = (float)(dcs * (double)sp.cs[j] * 0.5 * (1.0 - (double)t[kSlotWb + j] * t[kSlotWb + j]));
        }
        o[kSlotGamma] = (float)(dg * (double)sp.gamma * (double)kLn3 * (1.0 - (double)t[kSlotGamma] * t[kSlotGamma]));
        o[kSlotContrast] = (float)(dc * (1.0 - (double)t[kSlotContrast] * t[kSlotContrast]));
        o[kSlotUsm] = (float)(dp * 2.5 * (1.0 - (double)t[kSlotUsm] * t[kSlotUsm]));
    }
    cluster.sync();  // keep the peers' shared memory alive until CTA 0 has read it
}

constexpr size_t kBwdSmem = (size_t)(kXRingB * kXP + kHRing * kHP + kRB * kBP + 2 * kMaxU) * sizeof(float);

template <bool HAS_ICA, bool FAST, bool ALIGNED, bool TMA>
static int launch_bwd4(const CUtensorMap& gmap, const float* x, const float* A, const float* IcA, const float* feat, const float* g,
                       float* part, float* Spart, float* dx, int B, int H, int W, const Sched& sc, cudaStream_t st) {
    DD_ENSURE_SMEM((recovery_bwd_kernel<HAS_ICA, FAST, ALIGNED, TMA>), kBwdSmem, "recovery kernel");
    launch_pdl(recovery_bwd_kernel<HAS_ICA, FAST, ALIGNED, TMA>, dim3(sc.G), dim3(kThreads), kBwdSmem, st, gmap, x, A, IcA, feat, g, part,
               Spart, dx, B, H, W, (const float*)nullptr);
    return DD_OK;
}

// x is the uint8 batch, read through the darkening table
template <bool HAS_ICA, bool FAST>
static int launch_bwd_u8(const uint8_t* src, const float* tab, const float* A, const float* IcA, const float* feat, const float* g, float* dfeat,
                         int B, int H, int W, float* ws, cudaStream_t st) {
    const Sched sc = make_sched(B, H, W);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    const bool tma = W >= kXP && H >= 16 && make_tensor_map_3d(&gmap, g, B * 3, H, W, kXP, 16);
    const float* xf = reinterpret_cast<const float*>(src);
    if (tma) {
        auto kern = recovery_bwd_kernel<HAS_ICA, FAST, true, true, kXU8>;
        DD_ENSURE_SMEM(kern, kBwdSmem, "recovery kernel (uint8 source)");
        launch_pdl(kern, dim3(sc.G), dim3(kThreads), kBwdSmem, st, gmap, xf, A, IcA, feat, g, part, Spart, (float*)nullptr, B, H, W, tab);
    } else {
        auto kern = recovery_bwd_kernel<HAS_ICA, FAST, true, false, kXU8>;
        DD_ENSURE_SMEM(kern, kBwdSmem, "recovery kernel (uint8 source)");
        launch_pdl(kern, dim3(sc.G), dim3(kThreads), kBwdSmem, st, gmap, xf, A, IcA, feat, g, part, Spart, (float*)nullptr, B, H, W, tab);
    }
    launch_pdl_cluster(recovery_bwd_finalize_kernel<HAS_ICA, FAST, unsigned char>, dim3(B * kFinCluster), dim3(kFinThreads), 0, st, kFinCluster,
                       (const unsigned char*)src, A, IcA, feat, (const float*)part, (const float*)Spart, dfeat, (unsigned char*)nullptr, B, H, W, sc.G,
                       sc.strips, tab);
    count_launch(2);
    return check_launch("dd_recovery_bwd_u8");
}

// bf16 I/O mode on the CUDA-core kernel (no dx): x and / or g are bf16, fp32 arithmetic and FFMA2 blur.  Measured on B200
// (16x3x640x640, both bf16): this kernel ~135 us against 232 us for the tensor-core backward, whose lane <-> column epilogue pays
// one warp reduction per 32 pixels and scalar loads (profiles/r02_ncu_full_filters_bf16_tc.md) -- so it is the default.
template <bool HAS_ICA, typename TX, typename TG>
static int launch_bwd_cc_io(const TX* x, const float* A, const float* IcA, const float* feat, const TG* g, float* dfeat, int B, int H, int W, float* ws,
                            cudaStream_t st) {
    constexpr int XT = sizeof(TX) == 2 ? kXBF16 : kXF32;
    constexpr bool GB = sizeof(TG) == 2;
    const Sched sc = make_sched(B, H, W);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    const float* xf = reinterpret_cast<const float*>(x);
    const float* gf = reinterpret_cast<const float*>(g);
    bool tma = false;
    if (!GB) tma = W >= kXP && H >= 16 && make_tensor_map_3d(&gmap, gf, B * 3, H, W, kXP, 16);
    if (tma) {
        auto kern = recovery_bwd_kernel<HAS_ICA, true, true, !GB, XT, false>;
        DD_ENSURE_SMEM(kern, kBwdSmem, "recovery kernel (bf16 I/O)");
        launch_pdl(kern, dim3(sc.G), dim3(kThreads), kBwdSmem, st, gmap, xf, A, IcA, feat, gf, part, Spart, (float*)nullptr, B, H, W, (const float*)nullptr);
    } else {
        auto kern = recovery_bwd_kernel<HAS_ICA, true, true, false, XT, GB>;
        DD_ENSURE_SMEM(kern, kBwdSmem, "recovery kernel (bf16 I/O)");
        launch_pdl(kern, dim3(sc.G), dim3(kThreads), kBwdSmem, st, gmap, xf, A, IcA, feat, gf, part, Spart, (float*)nullptr, B, H, W, (const float*)nullptr);
    }
    launch_pdl_cluster(recovery_bwd_finalize_kernel<HAS_ICA, true, TX>, dim3(B * kFinCluster), dim3(kFinThreads), 0, st, kFinCluster, x, A, IcA, feat,
                       (const float*)part, (const float*)Spart, dfeat, (TX*)nullptr, B, H, W, sc.G, sc.strips, (const float*)nullptr);
    count_launch(2);
    return check_launch("dd_recovery_bwd_ex");
}

// tensor-core variant: 148 persistent CTAs, 3xTF32 (fp32 gates).  Same workspace, same finalize kernel.
template <bool HAS_ICA, bool FAST>
static int launch_bwd_tc(const float* x, const float* A, const float* IcA, const float* feat, const float* g, float* dfeat,
                         float* dx, int B, int H, int W, float* ws, cudaStream_t st) {
    constexpr int R = 48;
    constexpr bool X3 = true;
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    auto kern = btc::recovery_bwd_tc_kernel<R, X3, HAS_ICA, FAST>;
    constexpr size_t smem = btc::Lay<R, X3>::SMEM;
    DD_ENSURE_SMEM(kern, smem, "recovery_bwd_tc_kernel");
    launch_pdl(kern, dim3(sc.G), dim3(btc::kThreadsTC), smem, st, x, A, IcA, feat, g, part, Spart, dx, B, H, W);
    launch_pdl_cluster(recovery_bwd_finalize_kernel<HAS_ICA, FAST>, dim3(B * kFinCluster), dim3(kFinThreads), 0, st, kFinCluster, x, A, IcA, feat,
                       (const float*)part, (const float*)Spart, dfeat, dx, B, H, W, sc.G, sc.strips * btc::kSpartPerStrip, (const float*)nullptr);
    count_launch(2);
    return check_launch("dd_recovery_bwd (tensor-core blur)");
}

// bf16 I/O mode: x (and dx) and / or the cotangent are bf16; plain TF32 blur on the tensor cores; same finalize kernel, which
// re-reads columns 0..2 of x -- it takes the element type as a template parameter
template <bool HAS_ICA, typename TX, typename TG>
static int launch_bwd_tc_io(const TX* x, const float* A, const float* IcA, const float* feat, const TG* g, float* dfeat, TX* dx, int B, int H,
                            int W, float* ws, cudaStream_t st) {
    constexpr int R = 48;
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    auto kern = btc::recovery_bwd_tc_kernel<R, false, HAS_ICA, true, TX, TG>;
    constexpr size_t smem = btc::Lay<R, false>::SMEM;
    DD_ENSURE_SMEM(kern, smem, "recovery_bwd_tc_kernel (bf16 I/O)");
    launch_pdl(kern, dim3(sc.G), dim3(btc::kThreadsTC), smem, st, x, A, IcA, feat, g, part, Spart, dx, B, H, W);
    launch_pdl_cluster(recovery_bwd_finalize_kernel<HAS_ICA, true, TX>, dim3(B * kFinCluster), dim3(kFinThreads), 0, st, kFinCluster, x, A, IcA, feat,
                       (const float*)part, (const float*)Spart, dfeat, dx, B, H, W, sc.G, sc.strips * btc::kSpartPerStrip, (const float*)nullptr);
    count_launch(2);
    return check_launch("dd_recovery_bwd_ex");
}

template <bool HAS_ICA, bool FAST, bool ALIGNED>
static int launch_bwd3(const float* x, const float* A, const float* IcA, const float* feat, const float* g, float* dfeat,
                       float* dx, int B, int H, int W, float* ws, cudaStream_t st) {
    // (H or W == 13: the two image borders are exactly one blur radius apart and the doubled-border form of the reflect adjoint
    //  used by the tensor-core kernel does not hold -- dd_blur_tc.cuh -- so those shapes stay on the CUDA-core kernel)
    if (ALIGNED && H > kRadius + 1 && W > kRadius + 1 && blur_on_tensor_cores())
        return launch_bwd_tc<HAS_ICA, FAST>(x, A, IcA, feat, g, dfeat, dx, B, H, W, ws, st);
    const Sched sc = make_sched(B, H, W);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    CUtensorMap gmap;
    memset(&gmap, 0, sizeof(gmap));
    // TMA staging needs 16-byte aligned rows (ALIGNED) and an image at least one box large; otherwise cp.async / scalar
    const bool tma = ALIGNED && W >= kXP && H >= 16 && make_tensor_map_3d(&gmap, g, B * 3, H, W, kXP, 16);
    if (int e = tma ? launch_bwd4<HAS_ICA, FAST, ALIGNED, ALIGNED>(gmap, x, A, IcA, feat, g, part, Spart, dx, B, H, W, sc, st)
                    : launch_bwd4<HAS_ICA, FAST, ALIGNED, false>(gmap, x, A, IcA, feat, g, part, Spart, dx, B, H, W, sc, st))
        return e;
    launch_pdl_cluster(recovery_bwd_finalize_kernel<HAS_ICA, FAST>, dim3(B * kFinCluster), dim3(kFinThreads), 0, st, kFinCluster, x, A, IcA, feat,
                       (const float*)part, (const float*)Spart, dfeat, dx, B, H, W, sc.G, sc.strips, (const float*)nullptr);
    count_launch(2);
    return check_launch("dd_recovery_bwd");
}

template <bool HAS_ICA, bool FAST>
static int launch_bwd2(const float* x, const float* A, const float* IcA, const float* feat, const float* g, float* dfeat,
                       float* dx, int B, int H, int W, float* ws, cudaStream_t st) {
    const bool aligned = (W & 3) == 0 && (((uintptr_t)x | (uintptr_t)g | (uintptr_t)dx) & 15) == 0;
    return aligned ? launch_bwd3<HAS_ICA, FAST, true>(x, A, IcA, feat, g, dfeat, dx, B, H, W, ws, st)
                   : launch_bwd3<HAS_ICA, FAST, false>(x, A, IcA, feat, g, dfeat, dx, B, H, W, ws, st);
}

}  // namespace dd

extern "C" int dd_recovery_bwd(const float* x, const float* A, const float* IcA, const float* feat, const float* g_, float* dfeat, float* dx, int B,
                               int H, int W, void* ws, size_t ws_bytes, void* stream_);

extern "C" int dd_recovery_bwd_ex(const void* x, int x_dtype, const float* A, const float* IcA, const float* feat, const void* g_, int g_dtype,
                                  float* dfeat, void* dx, int B, int H, int W, void* ws, size_t ws_bytes, void* stream_) {
    using namespace dd;
    typedef __nv_bfloat16 bf16;
    if (x_dtype == DD_F32 && g_dtype == DD_F32)
        return dd_recovery_bwd(reinterpret_cast<const float*>(x), A, IcA, feat, reinterpret_cast<const float*>(g_), dfeat, reinterpret_cast<float*>(dx), B, H,
                               W, ws, ws_bytes, stream_);
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_bwd_ex", B, H, W)) return e;
    DD_REQUIRE((x_dtype == DD_F32 || x_dtype == DD_BF16) && (g_dtype == DD_F32 || g_dtype == DD_BF16), DD_ERR_INVALID, "dd_recovery_bwd_ex: unknown dtype");
    DD_REQUIRE(x && feat && g_ && dfeat, DD_ERR_INVALID, "dd_recovery_bwd_ex: null pointer");
    DD_REQUIRE(ws && ws_bytes >= recovery_bwd_ws_bytes(B, H, W), DD_ERR_WORKSPACE, "dd_recovery_bwd_ex: workspace %zu < %zu", ws_bytes,
               recovery_bwd_ws_bytes(B, H, W));
    DD_REQUIRE((W & 3) == 0 && H > kRadius + 1 && W > kRadius + 1 && (((uintptr_t)x | (uintptr_t)g_ | (uintptr_t)dx) & 15) == 0, DD_ERR_INVALID,
               "dd_recovery_bwd_ex: the bf16 I/O mode needs W %% 4 == 0, H and W >= 14 and 16-byte aligned x / g / dx (got %d x %d)", H, W);
    float* w = reinterpret_cast<float*>(ws);
    // no dx wanted (training: x is data) -> the CUDA-core kernel with bf16 loads (faster); DEDARK_BLUR=tc keeps the tensor-core backward
    if (!dx && !blur_on_tensor_cores()) {
#define DD_BWD_CC(TX, TG)                                                                                                                       \
    return IcA ? launch_bwd_cc_io<true, TX, TG>(reinterpret_cast<const TX*>(x), A, IcA, feat, reinterpret_cast<const TG*>(g_), dfeat, B, H, W, w, st) \
               : launch_bwd_cc_io<false, TX, TG>(reinterpret_cast<const TX*>(x), A, nullptr, feat, reinterpret_cast<const TG*>(g_), dfeat, B, H, W, w, st)
        if (x_dtype == DD_BF16 && g_dtype == DD_BF16) { DD_BWD_CC(bf16, bf16); }
        if (x_dtype == DD_BF16) { DD_BWD_CC(bf16, float); }
        DD_BWD_CC(float, bf16);
#undef DD_BWD_CC
    }
#define DD_BWD_IO(TX, TG)                                                                                                                       \
    return IcA ? launch_bwd_tc_io<true, TX, TG>(reinterpret_cast<const TX*>(x), A, IcA, feat, reinterpret_cast<const TG*>(g_), dfeat,              \
                                                reinterpret_cast<TX*>(dx), B, H, W, w, st)                                                        \
               : launch_bwd_tc_io<false, TX, TG>(reinterpret_cast<const TX*>(x), A, nullptr, feat, reinterpret_cast<const TG*>(g_), dfeat,         \
                                                 reinterpret_cast<TX*>(dx), B, H, W, w, st)
    if (x_dtype == DD_BF16 && g_dtype == DD_BF16) { DD_BWD_IO(bf16, bf16); }
    if (x_dtype == DD_BF16) { DD_BWD_IO(bf16, float); }
    DD_BWD_IO(float, bf16);
#undef DD_BWD_IO
}

extern "C" int dd_recovery_bwd(const float* x, const float* A, const float* IcA, const float* feat, const float* g_,
                               float* dfeat, float* dx, int B, int H, int W, void* ws, size_t ws_bytes,
                               void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_bwd", B, H, W)) return e;
    DD_REQUIRE(x && feat && g_ && dfeat, DD_ERR_INVALID, "dd_recovery_bwd: null pointer");
    DD_REQUIRE(ws && ws_bytes >= recovery_bwd_ws_bytes(B, H, W), DD_ERR_WORKSPACE,
               "dd_recovery_bwd: workspace %zu < %zu", ws_bytes, recovery_bwd_ws_bytes(B, H, W));
    const bool fast = !precise_mode();
    float* w = reinterpret_cast<float*>(ws);
    if (IcA)
        return fast ? launch_bwd2<true, true>(x, A, IcA, feat, g_, dfeat, dx, B, H, W, w, st)
                    : launch_bwd2<true, false>(x, A, IcA, feat, g_, dfeat, dx, B, H, W, w, st);
    return fast ? launch_bwd2<false, true>(x, A, nullptr, feat, g_, dfeat, dx, B, H, W, w, st)
                : launch_bwd2<false, false>(x, A, nullptr, feat, g_, dfeat, dx, B, H, W, w, st);
}

// ---- backward of the chain on the uint8 batch (no darkened fp32 batch in HBM) --------------------------
extern "C" int dd_recovery_bwd_u8(const uint8_t* src, const float* dark_table, const float* A, const float* IcA, const float* feat, const float* g_,
                                  float* dfeat, int B, int H, int W, void* ws, size_t ws_bytes, void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_bwd_u8", B, H, W)) return e;
    DD_REQUIRE(src && dark_table && feat && g_ && dfeat, DD_ERR_INVALID, "dd_recovery_bwd_u8: null pointer");
    DD_REQUIRE(ws && ws_bytes >= recovery_bwd_ws_bytes(B, H, W), DD_ERR_WORKSPACE, "dd_recovery_bwd_u8: workspace %zu < %zu", ws_bytes,
               recovery_bwd_ws_bytes(B, H, W));
    DD_REQUIRE((W & 3) == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)g_ & 15) == 0 && (!IcA || ((uintptr_t)IcA & 15) == 0), DD_ERR_INVALID,
               "dd_recovery_bwd_u8: needs W %% 4 == 0, a 4-byte aligned source and 16-byte aligned g / IcA (got W = %d)", W);
    const bool fast = !precise_mode();
    float* w = reinterpret_cast<float*>(ws);
    if (IcA)
        return fast ? launch_bwd_u8<true, true>(src, dark_table, A, IcA, feat, g_, dfeat, B, H, W, w, st)
                    : launch_bwd_u8<true, false>(src, dark_table, A, IcA, feat, g_, dfeat, B, H, W, w, st);
    return fast ? launch_bwd_u8<false, true>(src, dark_table, A, nullptr, feat, g_, dfeat, B, H, W, w, st)
                : launch_bwd_u8<false, false>(src, dark_table, A, nullptr, feat, g_, dfeat, B, H, W, w, st);
}
