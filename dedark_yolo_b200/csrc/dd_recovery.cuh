// dd_recovery.cuh -- pieces shared by the fused filter-chain kernels (dd_recovery_fwd.cu, dd_recovery_bwd.cu):
// per-image regressors, the pointwise chain, the two FFMA2 blur passes and the persistent-CTA work list.
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,295-304,316-317.
#pragma once
#include <cstdlib>

#include "dd_common.cuh"
#include "dd_layout.cuh"

namespace dd {

typedef unsigned long long u64;

constexpr int kThreads = 256;
constexpr int kRB = kRowBlock;                // 32 rows per marching block
constexpr int kXW = kStripW + 2 * kRadius;    // 152 staged columns
constexpr int kXW4 = kXW / 4;                 // 38 float4 per staged row
constexpr int kXP = 156;                      // XS pitch (floats): (kXP/4) odd -> conflict-free LDS.128 by row
constexpr int kHRing = 64;                    // HS ring depth (rows), power of two
constexpr int kHP = 132;                      // HS pitch (floats)
constexpr int kStage4 = (kRB * kXW4 + kThreads - 1) / kThreads;  // float4 staged per thread per block (5)
constexpr int kMaxU = kMaxSegRows + 2 * kRadius;

// ---- packed fp32x2 helpers (sm_100 FFMA2) -----------------------------------------------------------
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float lg2_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_fast(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__host__ __device__ __forceinline__ constexpr float tapj(int j) {  // tap of window position j in 0..24
    return tap(j < kRadius ? kRadius - j : j - kRadius);
}

// 16-byte async global->shared copy (LDGSTS); `valid == false` zero-fills the destination
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
// 4-byte variant (any alignment); used where the shared-memory layout interleaves rows
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- per-image parameters ----------------------------------------------------------------------------
struct ImgParams {
    float w, s[3], gamma, c, p;
    float t[kFeat];  // tanh of the raw features (needed by the Jacobians)
    float cs[3], Z;
};

// filtersB.py:151-152,186-187,227-229,246-256,296-297 + util_filters.py:295-304
__device__ inline void regress(const float* __restrict__ f, ImgParams& P) {
    for (int i = 0; i < kFeat; ++i) P.t[i] = 0.f;
    P.t[0] = tanhf(f[kSlotDedark]);
    P.w = P.t[0] * 0.9f / 2.0f + 0.55f;
    const float mask[3] = {0.f, 1.f, 1.f};
    for (int j = 0; j < 3; ++j) {
        P.t[kSlotWb + j] = tanhf(f[kSlotWb + j] * mask[j]);
        P.cs[j] = expf(P.t[kSlotWb + j] * 1.0f / 2.0f);
    }
    P.Z = kWbEps + kLumR * P.cs[0] + kLumG * P.cs[1] + kLumB * P.cs[2];
    for (int j = 0; j < 3; ++j) P.s[j] = P.cs[j] / P.Z;
    P.t[kSlotGamma] = tanhf(f[kSlotGamma]);
    P.gamma = expf(P.t[kSlotGamma] * kLn3);
    P.t[kSlotContrast] = tanhf(f[kSlotContrast]);
    P.c = P.t[kSlotContrast];
    P.t[kSlotUsm] = tanhf(f[kSlotUsm]);
    P.p = P.t[kSlotUsm] * 5.0f / 2.0f + 2.5f;
}

// The same regressors computed by one WARP (all 32 lanes call it): the seven tanh and the four exp run in two parallel
// levels instead of eleven dependent libm calls on one thread while the whole CTA waits (the values are bit-identical:
// same expressions, same order of the final sums).
__device__ inline void regress_warp(const float* __restrict__ f, ImgParams& P) {
    const int lane = threadIdx.x & 31;
    if (lane < kFeat) {
        const bool wb = lane >= kSlotWb && lane < kSlotWb + 3;
        const bool live = lane == kSlotDedark || wb || lane == kSlotGamma || lane == kSlotContrast || lane == kSlotUsm;
        float t = 0.f;
        const float fv = live ? __ldcg(f + lane) : 0.f;   // the predictor's output: ld.global.cg, never moved across griddepcontrol.wait
        if (live) t = tanhf(wb ? fv * (lane == kSlotWb ? 0.f : 1.f) : fv);
        P.t[lane] = t;
        if (wb) P.cs[lane - kSlotWb] = expf(t * 1.0f / 2.0f);
        if (lane == kSlotGamma) P.gamma = expf(t * kLn3);
    }
    __syncwarp();
    if (lane == 0) {
        P.w = P.t[kSlotDedark] * 0.9f / 2.0f + 0.55f;
        P.Z = kWbEps + kLumR * P.cs[0] + kLumG * P.cs[1] + kLumB * P.cs[2];
        for (int j = 0; j < 3; ++j) P.s[j] = P.cs[j] / P.Z;
        P.c = P.t[kSlotContrast];
        P.p = P.t[kSlotUsm] * 5.0f / 2.0f + 2.5f;
    }
}

// Per-(image, channel) constants of the pointwise chain.  With the default IcA (a constant 0.5) DeDark and WB
// collapse into one FMA:  x2 = x0 * k1 + k0,  k1 = s / txc,  k0 = (a - a / txc) * s.
struct ChainK {
    float w, a, s, gamma, k1, k0, inv;  // inv = 1 / txc (default IcA only)
};
__device__ __forceinline__ ChainK make_chain(const ImgParams& P, int ch, float a) {
    ChainK k;
    k.w = P.w; k.a = a; k.s = P.s[ch]; k.gamma = P.gamma;
    const float txc = fmaxf(1.f - P.w * kDefaultIcA, kTxMin);
    k.inv = 1.f / txc;
    k.k1 = k.inv * k.s;
    k.k0 = (a - a * k.inv) * k.s;
    return k;
}

// DeDark (filtersB.py:211-214) -> WB (:259): returns x2 (before the 1e-4 clamp)
template <bool HAS_ICA>
__device__ __forceinline__ float chain_x2(const ChainK& k, float x0, float ica) {
    if (HAS_ICA) {
        const float txc = fmaxf(fmaf(-k.w, ica, 1.f), kTxMin);
        return (__fdiv_rn(x0 - k.a, txc) + k.a) * k.s;  // caller-supplied IcA: tx may sit at 0.01 (x100 gain), stay IEEE
    }
    return fmaf(x0, k.k1, k.k0);
}
// Gamma (filtersB.py:232-233): pow(max(x2, 1e-4), gamma); *lg receives log2 of the clamped base.
// FAST: ex2(gamma * lg2(x)) on the MUFU pipe (max error 8e-7 relative for gamma <= 1, 3.6e-6 at gamma = 3,
// profiles/microbench/fastpow.cu); otherwise powf/log2f.
template <bool FAST>
__device__ __forceinline__ float gamma_pow(float x2c, float gamma, float* lg) {
    if (FAST) {
        const float l = lg2_fast(x2c);
        if (lg) *lg = l;
        return ex2_fast(gamma * l);
    }
    if (lg) *lg = log2f(x2c);
    return powf(x2c, gamma);
}
template <bool HAS_ICA, bool FAST>
__device__ __forceinline__ float chain_x3(const ChainK& k, float x0, float ica) {
    return gamma_pow<FAST>(fmaxf(chain_x2<HAS_ICA>(k, x0, ica), kGammaClamp), k.gamma, nullptr);
}

struct RowLum {
    float lraw, lum, cl, denom, q;
};
// the rgb2lum quirk (util_filters.py:270-273 on an NCHW tensor): per (b, ch, row), from x3 at columns 0..2.
// LITERAL = the reference's expression -cos(pi lum)*0.5+0.5 (filtersB.py:301), used by the forward so that its
// fp32 rounding matches the reference's; it cancels badly for lum << 1, so the backward (graded against the fp64
// truth) evaluates the same quantity as sin^2(pi lum / 2).
template <bool LITERAL>
__device__ __forceinline__ RowLum row_lum(float x3_0, float x3_1, float x3_2) {
    RowLum r;
    r.lraw = kLumR * x3_0 + kLumG * x3_1 + kLumB * x3_2;
    r.lum = fminf(fmaxf(r.lraw, 0.f), 1.f);
    if (LITERAL) {
        r.cl = -cosf(kPi * r.lum) * 0.5f + 0.5f;
    } else {
        const float sn = sinf(0.5f * kPi * r.lum);
        r.cl = sn * sn;
    }
    r.denom = r.lum + kContrastEps;
    r.q = r.cl / r.denom;
    return r;
}

// ---- the two blur passes ------------------------------------------------------------------------------
// 8 horizontally adjacent outputs from 32 staged values; output t uses in[t .. t+24].  Outputs are computed in
// pairs (t, t+1) with ONE tap broadcast to both lanes (an FFMA2 immediate): even window positions use the
// aligned input pairs E[i] = (in[2i], in[2i+1]) exactly as LDS.128 delivered them, odd positions use the
// shifted pairs O[i] = (in[2i+1], in[2i+2]) (two register moves each, on the ALU pipe).
__device__ __forceinline__ void hpass8(const float* __restrict__ xrow, float out[8]) {
    u64 E[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xrow + 4 * i);
        E[2 * i] = v.x;
        E[2 * i + 1] = v.y;
    }
    // even window positions: the aligned input pairs E[i] = (in[2i], in[2i+1]) exactly as LDS.128 delivered them, one FFMA2
    // with an immediate tap.  Odd positions would need the shifted pairs (in[2i+1], in[2i+2]) -- two register moves each, ~80
    // moves per call once the compiler has satisfied the pair alignment -- so they are two scalar FFMAs on the halves
    // instead: the same FMA-pipe time, a quarter fewer instructions.  (Round 2 tried the all-FFMA2 form -- odd taps on the aligned
    // pairs into a swapped accumulator (-> out[t+1], -> out[t]) with tap pairs (k_{m-1}, k_{m+1}) from uniform registers, 26 FFMA2
    // per output pair and bit-identical sums: 135.3 -> 140.3 us for the backward.  These passes are bound by the FMA pipe, not by
    // issue slots, and the packed form spends two more lane-FMAs per pair on zero taps.)
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
        u64 acc = pk(0.f, 0.f);
        float ox = 0.f, oy = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            if (j & 1) {
                ox = fmaf(tapj(j), upk(E[(t + j - 1) / 2]).y, ox);
                oy = fmaf(tapj(j), upk(E[(t + j + 1) / 2]).x, oy);
            } else {
                acc = fma2(E[(t + j) / 2], pk(tapj(j), tapj(j)), acc);
            }
        }
        const float2 a = upk(acc);
        out[t] = a.x + ox;
        out[t + 1] = a.y + oy;
    }
}

// 2 adjacent columns x 8 vertically adjacent outputs; Q[i] = ring row (first - 12 + i), columns (col2, col2+1)
__device__ __forceinline__ void vpass8x2(const float* __restrict__ HS, int base_slot, int col2, u64 out[8]) {
    // scatter form: each staged row is consumed as soon as it arrives (few live registers, 8 independent FMA chains).
    // base_slot is a multiple of 8 and the ring depth is too: a group of 8 rows never wraps, so four group pointers
    // (computed once) and immediate offsets replace 32 per-load ring computations.
#pragma unroll
    for (int r = 0; r < 8; ++r) out[r] = pk(0.f, 0.f);
    const float* grp[4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) grp[gi] = HS + ((base_slot + 8 * gi) & (kHRing - 1)) * kHP + col2;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const u64 q = *reinterpret_cast<const u64*>(grp[i >> 3] + (i & 7) * kHP);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int j = i - r;
            if (j >= 0 && j < kTaps) out[r] = fma2(q, pk(tapj(j), tapj(j)), out[r]);
        }
    }
}

// Generic 25-tap pass on packed pairs: 32 consecutive 64-bit pairs at p, p + stride, ... (stride in floats, a
// compile-time constant so every load carries an immediate offset) -> 8 pair outputs, out[t] = sum_j k[j] P[t+j].
// Both lanes of a pair share the tap (FFMA2 immediate).  Scatter form: a pair is consumed as soon as it is loaded.
//   H pass of the forward: pairs = (row 2r, row 2r+1) of one column in the row-pair-interleaved ring, STRIDE = 2
//   V pass:                pairs = (col 2c, col 2c+1) of one row in the row-major ring,               STRIDE = kHP
template <int STRIDE>
__device__ __forceinline__ void blur8_pairs(const float* __restrict__ p, u64 out[8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) out[r] = pk(0.f, 0.f);
    if (STRIDE == 2) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p + 2 * i);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (i - r >= 0 && i - r < kTaps) out[r] = fma2(q.x, pk(tapj(i - r), tapj(i - r)), out[r]);
                if (i + 1 - r >= 0 && i + 1 - r < kTaps) out[r] = fma2(q.y, pk(tapj(i + 1 - r), tapj(i + 1 - r)), out[r]);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const u64 q = *reinterpret_cast<const u64*>(p + i * STRIDE);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (i - r >= 0 && i - r < kTaps) out[r] = fma2(q, pk(tapj(i - r), tapj(i - r)), out[r]);
        }
    }
}

// ---- the work list of one CTA -----------------------------------------------------------------------------
struct Seg {
    int ps, plane, b, ch, strip, r0, r1, c0, nU, nB, seg_len;
};
// next run of row-blocks [blk, ...) inside one plane-strip, capped at kMaxSegRows rows and at blk_end
__device__ __forceinline__ Seg next_seg(long long blk, long long blk_end, const Sched& sc, int H) {
    Seg s;
    s.ps = (int)(blk / sc.nRB);
    const int rb0 = (int)(blk - (long long)s.ps * sc.nRB);
    long long n = blk_end - blk;
    if (n > sc.nRB - rb0) n = sc.nRB - rb0;
    if (n > kMaxSegRows / kRB) n = kMaxSegRows / kRB;
    s.plane = s.ps / sc.strips;
    s.strip = s.ps - s.plane * sc.strips;
    s.b = s.plane / 3;
    s.ch = s.plane - 3 * s.b;
    s.r0 = rb0 * kRB;
    s.r1 = min(H, (rb0 + (int)n) * kRB);
    s.c0 = s.strip * kStripW;
    s.seg_len = s.r1 - s.r0;
    s.nU = s.seg_len + 2 * kRadius;
    s.nB = (s.nU + kRB - 1) / kRB;
    return s;
}
__device__ __forceinline__ int seg_blocks(const Seg& s) { return (s.seg_len + kRB - 1) / kRB; }

// ---- host helpers -----------------------------------------------------------------------------------------
inline int check_recovery_shape(const char* who, int B, int H, int W) {
    DD_REQUIRE(B > 0 && H > 0 && W > 0, DD_ERR_INVALID, "%s: B, H, W must be positive (got %d, %d, %d)", who, B, H, W);
    DD_REQUIRE(W >= 3, DD_ERR_WIDTH_LT3, "%s: W = %d < 3: rgb2lum indexes columns 0..2", who, W);
    DD_REQUIRE(H > kRadius && W > kRadius, DD_ERR_REFLECT_PAD,
               "%s: reflect padding of 12 needs H, W > 12 (got %d x %d)", who, H, W);
    return DD_OK;
}

template <typename K>
inline int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return DD_ERR_CUDA;
    }
    return DD_OK;
}

// Which blur engine the fp32 entry points use.  DEDARK_BLUR=tc: the tensor-core kernels (dd_blur_tc.cuh, 3xTF32) whenever the
// rows are 16-byte aligned (W % 4 == 0); DEDARK_BLUR=cc or unset: the CUDA-core (FFMA2) kernels, which are the faster ones at
// fp32 precision (B200, 16x3x640x640, kernel time: forward 82 us vs 115 us, backward 135 us vs 256 us -- DESIGN.md section 4.5).
// The bf16 I/O mode runs its forward on the tensor cores (plain TF32, 81 us) and its backward on the CUDA-core kernel with bf16
// loads (151 us; 232 us on the tensor cores, which DEDARK_BLUR=tc and a requested dx still select).  Read on every call so that
// tests can compare the engines in one process.
inline bool blur_on_tensor_cores() {
    const char* e = getenv("DEDARK_BLUR");
    return e && e[0] == 't' && e[1] == 'c';
}

inline bool precise_mode() {  // DEDARK_PRECISE=1: powf/log2f instead of the MUFU pow (debugging aid, read once)
    static const bool v = [] { const char* e = getenv("DEDARK_PRECISE"); return e && e[0] == '1'; }();
    return v;
}

}  // namespace dd
