// dd_recovery.cu -- a6..a12 + a14: the fused filter chain (regressors, DeDark -> WB -> Gamma ->
// Contrast -> USM), forward and backward, one pass over the image each.
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,
// 295-304,316-317.  Closed-form backward: SURVEY.md section 8(a) row a14 (checked against autograd).
//
// Work decomposition ("marching strips", dd_layout.cuh).  Every (plane, 128-column strip) is cut into
// row-blocks of 32 rows; 296 persistent CTAs (2 per SM, 256 threads) each own a contiguous, equal
// range of row-blocks and march down them:
//     1. stage:  x0 -> pointwise chain -> x4, 32 rows x (128+24) columns, 128-bit global loads issued
//                one block ahead (register prefetch)                              -> smem ring XS
//     2. H pass: 25 taps, 8 outputs per thread from 32 staged values (8 LDS.128), packed FFMA2 with two
//                partial sums per output (adjacent inputs x adjacent taps)        -> smem ring HS
//     3. V pass: 25 taps, 2 columns x 8 rows per thread (32 LDS.64), FFMA2 with the tap broadcast to both
//                lanes, then y = (x4 - blur) p + x4 -> 64-bit coalesced global stores
// The vertical halo (24 rows) is paid once per CTA range, the pointwise chain is recomputed only on the
// 24 halo columns of a strip.  The rings hold 64 rows; XS pitch 156 / HS pitch 132 floats make the
// 128-bit shared accesses of pass 2 conflict-free (lanes map to rows).  The blur is FMA-pipe bound
// (50 FMA per pixel and pass pair ~ 28 us per 16x3x640^2 at 35 TFMA/s): FFMA2 halves its issue slots
// so loads, MUFU and index arithmetic issue in the shadow of the FMA pipe.
//
// pow(x, gamma) is evaluated as ex2(gamma * lg2(x)) on the MUFU pipe (measured max error 8e-7 relative
// for gamma <= 1, 3.6e-6 at gamma = 3: profiles/microbench/fastpow.cu); DEDARK_PRECISE=1 selects powf/logf.
//
// Backward uses the same skeleton on g = dL/dy: the ring holds g (zero outside the image), the two
// passes apply the ADJOINT of the reflect-padded blur (zero-padded correlation + the reflect
// fold-back as a gather: W(i,j) = k[j-i] + [j>=1] k[j+i] + [j<=n-2] k[2(n-1)-j-i]); the chain is
// recomputed from x and the parameter gradients are reduced per thread -> per (CTA, plane-strip)
// partials -> a fixed-order finalize kernel (no float atomics), which also applies the row-coupled
// fix-up of columns 0..2 (the rgb2lum quirk) and the regressor Jacobians.
#include <cstdlib>

#include "dd_common.cuh"
#include "dd_layout.cuh"

namespace dd {

typedef unsigned long long u64;

constexpr int kThreads = 256;
constexpr int kRB = kRowBlock;                // 32 rows per marching block
constexpr int kRing = 64;                     // ring depth (rows), power of two
constexpr int kXW = kStripW + 2 * kRadius;    // 152 staged columns
constexpr int kXW4 = kXW / 4;                 // 38 float4 per staged row
constexpr int kXP = 156;                      // XS pitch (floats): (kXP/4) odd -> conflict-free LDS.128 by row
constexpr int kHP = 132;                      // HS pitch (floats)
constexpr int kStage4 = (kRB * kXW4 + kThreads - 1) / kThreads;  // float4 staged per thread per block (5)
constexpr int kMaxU = kMaxSegRows + 2 * kRadius;
constexpr int kSP = 68;                       // SS pitch (floats)

__constant__ float c_tap[13] = {DD_K0, DD_K1, DD_K2, DD_K3, DD_K4, DD_K5, DD_K6,
                                DD_K7, DD_K8, DD_K9, DD_K10, DD_K11, DD_K12};

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

// ---- per-image parameters ----------------------------------------------------------------------------
struct ImgParams {
    float w, s[3], gamma, c, p;
    float t[kFeat];  // tanh of the raw features (needed by the Jacobians)
    float cs[3], Z;
};

// filtersB.py:151-152,186-187,227-229,246-256,296-297 + util_filters.py:295-304
__device__ void regress(const float* __restrict__ f, ImgParams& P) {
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

// DeDark (filtersB.py:211-214) -> WB (:259) -> clamp: returns x2 (before the 1e-4 clamp)
template <bool HAS_ICA>
__device__ __forceinline__ float chain_x2(const ChainK& k, float x0, float ica) {
    if (HAS_ICA) {
        const float txc = fmaxf(fmaf(-k.w, ica, 1.f), kTxMin);
        return (__fdiv_rn(x0 - k.a, txc) + k.a) * k.s;  // caller-supplied IcA: tx may sit at 0.01 (x100 gain), stay IEEE
    }
    return fmaf(x0, k.k1, k.k0);
}
// Gamma (filtersB.py:232-233): pow(max(x2, 1e-4), gamma); *lg receives log2 of the clamped base
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
// 8 horizontally adjacent outputs from 32 staged values; output t uses in[t .. t+24].  Inputs stay in the
// 64-bit pairs the LDS.128 delivered; each output keeps two partial sums (lane 0: even window positions,
// lane 1: odd ones) plus one scalar FMA for the unpaired tap.
__device__ __forceinline__ void hpass8(const float* __restrict__ xrow, float out[8]) {
    u64 P[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(xrow + 4 * i);
        P[2 * i] = v.x;
        P[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        u64 acc = pk(0.f, 0.f);
        if ((t & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 24; j += 2) acc = fma2(P[(t + j) / 2], pk(tapj(j), tapj(j + 1)), acc);
            const float2 a = upk(acc);
            out[t] = fmaf(upk(P[(t + 24) / 2]).x, tapj(24), a.x) + a.y;
        } else {
#pragma unroll
            for (int j = 1; j < 25; j += 2) acc = fma2(P[(t + j) / 2], pk(tapj(j), tapj(j + 1)), acc);
            const float2 a = upk(acc);
            out[t] = fmaf(upk(P[t / 2]).y, tapj(0), a.x) + a.y;
        }
    }
}

// 2 adjacent columns x 8 vertically adjacent outputs; Q[i] = ring row (first - 12 + i), columns (col2, col2+1)
__device__ __forceinline__ void vpass8x2(const float* __restrict__ HS, int base_slot, int col2, u64 out[8]) {
    u64 Q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i)
        Q[i] = *reinterpret_cast<const u64*>(HS + ((base_slot + i) & (kRing - 1)) * kHP + col2);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        u64 acc = pk(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kTaps; ++j) acc = fma2(Q[r + j], pk(tapj(j), tapj(j)), acc);
        out[r] = acc;
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

// =================================================================================================
// forward
// =================================================================================================
template <bool HAS_ICA, bool FAST>
__global__ void __launch_bounds__(kThreads, 2)
recovery_fwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                    const float* __restrict__ feat, float* __restrict__ y, int B, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* XS = smem;
    float* HS = XS + kRing * kXP;
    float* MS = HS + kRing * kHP;  // per virtual row: m = (1-c) + c*q
    __shared__ ImgParams sp;

    const int tid = threadIdx.x;
    const Sched sc = make_sched(B, H, W);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const bool w4 = (W & 3) == 0, w2 = (W & 1) == 0;

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        __syncthreads();  // previous segment fully consumed (rings, MS, sp)
        if (tid == 0) regress(feat + u.b * kFeat, sp);
        __syncthreads();
        const ChainK ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
        const float pc = sp.c, pp = sp.p;
        const float* xp = x + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
        float* yp = y + (size_t)u.plane * H * W;

        for (int v = tid; v < u.nU; v += kThreads) {
            const int row = reflect(u.r0 - kRadius + v, H);
            float x3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float ica = HAS_ICA ? __ldg(ip + (size_t)row * W + k) : kDefaultIcA;
                x3[k] = chain_x3<HAS_ICA, FAST>(ck, __ldg(xp + (size_t)row * W + k), ica);
            }
            const RowLum rl = row_lum<true>(x3[0], x3[1], x3[2]);
            MS[v] = (1.f - pc) + pc * rl.q;
        }

        float4 pre[kStage4], prei[kStage4];
        auto stage = [&](int n) {
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4;
                const int v = n * kRB + rr;
                if (f < kRB * kXW4 && v < u.nU) {
                    const int row = reflect(u.r0 - kRadius + v, H);
                    const int gc = u.c0 - kRadius + 4 * c4;
                    const size_t ro = (size_t)row * W;
                    if (w4 && gc >= 0 && gc + 3 < W) {
                        pre[k] = __ldg(reinterpret_cast<const float4*>(xp + ro + gc));
                        if (HAS_ICA) prei[k] = __ldg(reinterpret_cast<const float4*>(ip + ro + gc));
                    } else {
                        int g[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) g[e] = min(max(reflect(gc + e, W), 0), W - 1);
                        pre[k] = make_float4(__ldg(xp + ro + g[0]), __ldg(xp + ro + g[1]), __ldg(xp + ro + g[2]), __ldg(xp + ro + g[3]));
                        if (HAS_ICA)
                            prei[k] = make_float4(__ldg(ip + ro + g[0]), __ldg(ip + ro + g[1]), __ldg(ip + ro + g[2]), __ldg(ip + ro + g[3]));
                    }
                }
            }
        };
        stage(0);

        for (int n = 0; n < u.nB; ++n) {
            __syncthreads();  // MS ready (n == 0); ring slots of block n no longer read by the previous V pass
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4;
                const int v = n * kRB + rr;
                if (f < kRB * kXW4 && v < u.nU) {
                    const float m = MS[v];
                    const float4 in = pre[k];
                    const float4 ic = HAS_ICA ? prei[k] : make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                    float4 o;
                    o.x = chain_x3<HAS_ICA, FAST>(ck, in.x, ic.x) * m;
                    o.y = chain_x3<HAS_ICA, FAST>(ck, in.y, ic.y) * m;
                    o.z = chain_x3<HAS_ICA, FAST>(ck, in.z, ic.z) * m;
                    o.w = chain_x3<HAS_ICA, FAST>(ck, in.w, ic.w) * m;
                    *reinterpret_cast<float4*>(XS + (v & (kRing - 1)) * kXP + 4 * c4) = o;
                }
            }
            if (n + 1 < u.nB) stage(n + 1);
            __syncthreads();
            {   // horizontal pass: lanes -> rows, each thread two groups of 8 columns
                const int rr = tid & 31, cg = tid >> 5;
                const int slot = (n * kRB + rr) & (kRing - 1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cb = 64 * half + 8 * cg;
                    float o[8];
                    hpass8(XS + slot * kXP + cb, o);
                    float4* dst = reinterpret_cast<float4*>(HS + slot * kHP + cb);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
            __syncthreads();
            {   // vertical pass + USM epilogue: lanes -> column pairs, 8 rows per thread
                const int col2 = 2 * (tid & 63), rg = tid >> 6;
                const int o_first = n * kRB - kRadius + 8 * rg;  // virtual row of the first output
                if (o_first >= kRadius && o_first < kRadius + u.seg_len) {
                    u64 bl[8];
                    vpass8x2(HS, (o_first - kRadius) & (kRing - 1), col2, bl);
                    const int gc = u.c0 + col2;
                    const u64 p2 = pk(pp, pp), m1 = pk(-1.f, -1.f);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int o = o_first + r;
                        if (o < kRadius + u.seg_len && gc < W) {
                            const u64 x4 = *reinterpret_cast<const u64*>(XS + (o & (kRing - 1)) * kXP + col2 + kRadius);
                            const u64 yv = fma2(fma2(bl[r], m1, x4), p2, x4);  // (x4 - blur) * p + x4
                            float* dst = yp + (size_t)(u.r0 + o - kRadius) * W + gc;
                            if (w2) {
                                *reinterpret_cast<u64*>(dst) = yv;
                            } else {
                                const float2 t = upk(yv);
                                dst[0] = t.x;
                                if (gc + 1 < W) dst[1] = t.y;
                            }
                        }
                    }
                }
            }
        }
    }
}

// =================================================================================================
// backward
// =================================================================================================
template <bool HAS_ICA, bool FAST>
__global__ void __launch_bounds__(kThreads, 2)
recovery_bwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                    const float* __restrict__ feat, const float* __restrict__ g, float* __restrict__ part,
                    float* __restrict__ Spart, float* __restrict__ dx, int B, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* XS = smem;                 // g, zero outside the image
    float* HS = XS + kRing * kXP;
    float* MSm = HS + kRing * kHP;    // per virtual row: m = (1-c) + c*q
    float* MSq = MSm + kMaxU;         // per virtual row: q - 1
    float* SS = MSq + kMaxU;          // [32 rows][64 column pairs] partial row sums of g4 * x3
    __shared__ ImgParams sp;
    __shared__ float s_red[32];

    const int tid = threadIdx.x;
    const Sched sc = make_sched(B, H, W);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const bool w4 = (W & 3) == 0, w2 = (W & 1) == 0;

    float acc_p = 0.f, acc_c = 0.f, acc_g = 0.f, acc_s = 0.f, acc_w = 0.f;
    int cur_ps = -1;
    auto flush = [&]() {  // per (CTA, plane-strip) partial sums -> slot (cta + ps)
        float* out = part + (size_t)(blockIdx.x + cur_ps) * kBwdSums;
        float s;
        s = block_sum<float>(acc_p, s_red); if (tid == 0) out[0] = s;
        s = block_sum<float>(acc_c, s_red); if (tid == 0) out[1] = s;
        s = block_sum<float>(acc_g, s_red); if (tid == 0) out[2] = s;
        s = block_sum<float>(acc_s, s_red); if (tid == 0) out[3] = s;
        s = block_sum<float>(acc_w, s_red); if (tid == 0) out[4] = s;
        acc_p = acc_c = acc_g = acc_s = acc_w = 0.f;
    };

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        if (cur_ps >= 0 && u.ps != cur_ps) flush();
        cur_ps = u.ps;
        __syncthreads();
        if (tid == 0) regress(feat + u.b * kFeat, sp);
        __syncthreads();
        const float a = A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA;
        const ChainK ck = make_chain(sp, u.ch, a);
        const float pc = sp.c, pp = sp.p, pg = sp.gamma, ps = ck.s;
        const float* xp = x + (size_t)u.plane * H * W;
        const float* gp = g + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;

        for (int v = tid; v < u.nU; v += kThreads) {
            const int row = u.r0 - kRadius + v;
            float m = 0.f, q1 = 0.f;
            if (row >= 0 && row < H) {
                float x3[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float ica = HAS_ICA ? __ldg(ip + (size_t)row * W + k) : kDefaultIcA;
                    x3[k] = chain_x3<HAS_ICA, FAST>(ck, __ldg(xp + (size_t)row * W + k), ica);
                }
                const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
                q1 = rl.q - 1.f;
                m = (1.f - pc) + pc * rl.q;
            }
            MSm[v] = m;
            MSq[v] = q1;
        }

        float4 pre[kStage4];
        auto stage = [&](int n) {
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4;
                const int row = u.r0 - kRadius + n * kRB + rr;
                const int gc = u.c0 - kRadius + 4 * c4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (f < kRB * kXW4 && row >= 0 && row < H) {
                    const float* rp = gp + (size_t)row * W;
                    if (w4 && gc >= 0 && gc + 3 < W) {
                        v = __ldg(reinterpret_cast<const float4*>(rp + gc));
                    } else {
                        if (gc >= 0 && gc < W) v.x = __ldg(rp + gc);
                        if (gc + 1 >= 0 && gc + 1 < W) v.y = __ldg(rp + gc + 1);
                        if (gc + 2 >= 0 && gc + 2 < W) v.z = __ldg(rp + gc + 2);
                        if (gc + 3 >= 0 && gc + 3 < W) v.w = __ldg(rp + gc + 3);
                    }
                }
                pre[k] = v;
            }
        };
        stage(0);

        for (int n = 0; n < u.nB; ++n) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kStage4; ++k) {
                const int f = tid + k * kThreads;
                const int rr = f / kXW4, c4 = f - rr * kXW4;
                if (f < kRB * kXW4) *reinterpret_cast<float4*>(XS + ((n * kRB + rr) & (kRing - 1)) * kXP + 4 * c4) = pre[k];
            }
            if (n + 1 < u.nB) stage(n + 1);
            __syncthreads();
            {   // horizontal adjoint pass
                const int rr = tid & 31, cg = tid >> 5;
                const int slot = (n * kRB + rr) & (kRing - 1);
                const float* xrow = XS + slot * kXP;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cb = 64 * half + 8 * cg;
                    float o[8];
                    hpass8(xrow + cb, o);
                    const int j0 = u.c0 + cb;  // global column of o[0]
                    if (j0 <= kRadius || j0 + 7 >= W - 1 - kRadius) {  // reflect fold-back (image borders only)
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int j = j0 + t;
                            if (j >= 1 && j <= kRadius)
                                for (int i = 0; i <= kRadius - j; ++i) o[t] = fmaf(xrow[i - u.c0 + kRadius], c_tap[j + i], o[t]);
                            if (j >= W - 1 - kRadius && j <= W - 2)
                                for (int i = max(0, 2 * (W - 1) - j - kRadius); i <= W - 1; ++i)
                                    o[t] = fmaf(xrow[i - u.c0 + kRadius], c_tap[abs(2 * (W - 1) - j - i)], o[t]);
                        }
                    }
                    float4* dst = reinterpret_cast<float4*>(HS + slot * kHP + cb);
                    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
            __syncthreads();
            {   // vertical adjoint pass + chain recompute + reductions
                const int cp = tid & 63, col2 = 2 * cp, rg = tid >> 6;
                const int o_first = n * kRB - kRadius + 8 * rg;
                const bool group_ok = o_first >= kRadius && o_first < kRadius + u.seg_len;
                const int gc = u.c0 + col2;
                if (group_ok) {
                    u64 bt2[8];
                    vpass8x2(HS, (o_first - kRadius) & (kRing - 1), col2, bt2);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int o = o_first + r;
                        const int jr = u.r0 + o - kRadius;  // image row
                        float srow = 0.f;
                        if (o < kRadius + u.seg_len && gc < W) {
                            float2 bt = upk(bt2[r]);
                            if (jr >= 1 && jr <= kRadius)
                                for (int i = 0; i <= kRadius - jr; ++i) {
                                    const float2 h = *reinterpret_cast<const float2*>(HS + ((i - u.r0 + kRadius) & (kRing - 1)) * kHP + col2);
                                    bt.x = fmaf(h.x, c_tap[jr + i], bt.x);
                                    bt.y = fmaf(h.y, c_tap[jr + i], bt.y);
                                }
                            if (jr >= H - 1 - kRadius && jr <= H - 2)
                                for (int i = max(0, 2 * (H - 1) - jr - kRadius); i <= H - 1; ++i) {
                                    const float2 h = *reinterpret_cast<const float2*>(HS + ((i - u.r0 + kRadius) & (kRing - 1)) * kHP + col2);
                                    const float kk = c_tap[abs(2 * (H - 1) - jr - i)];
                                    bt.x = fmaf(h.x, kk, bt.x);
                                    bt.y = fmaf(h.y, kk, bt.y);
                                }
                            const float2 g5 = *reinterpret_cast<const float2*>(XS + (o & (kRing - 1)) * kXP + col2 + kRadius);
                            const size_t off = (size_t)jr * W + gc;
                            const bool two = gc + 1 < W;
                            float2 x0, ica = make_float2(kDefaultIcA, kDefaultIcA);
                            if (w2) {
                                x0 = __ldg(reinterpret_cast<const float2*>(xp + off));
                                if (HAS_ICA) ica = __ldg(reinterpret_cast<const float2*>(ip + off));
                            } else {
                                x0.x = __ldg(xp + off);
                                x0.y = two ? __ldg(xp + off + 1) : 0.f;
                                if (HAS_ICA) { ica.x = __ldg(ip + off); ica.y = two ? __ldg(ip + off + 1) : kDefaultIcA; }
                            }
                            const float m = MSm[o], q1 = MSq[o];
                            float dxv[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const float x0e = e ? x0.y : x0.x, icae = e ? ica.y : ica.x;
                                const float g5e = e ? g5.y : g5.x, bte = e ? bt.y : bt.x;
                                float txc = 0.f, inv;
                                bool pass_tx = true;
                                if (HAS_ICA) {
                                    const float tx = fmaf(-ck.w, icae, 1.f);
                                    pass_tx = tx >= kTxMin;
                                    txc = fmaxf(tx, kTxMin);
                                    inv = __fdiv_rn(1.f, txc);
                                } else {
                                    inv = ck.inv;
                                }
                                const float xa = x0e - a;
                                const float x1 = HAS_ICA ? __fdiv_rn(xa, txc) + a : fmaf(xa, inv, a);
                                const float x2 = x1 * ps;
                                const float x2c = fmaxf(x2, kGammaClamp);
                                float l2;
                                const float x3 = gamma_pow<FAST>(x2c, pg, &l2);
                                const float live = (e == 0 || two) ? 1.f : 0.f;
                                const float g4 = (g5e * (1.f + pp) - pp * bte) * live;
                                acc_p = fmaf(x3 * m, (g5e - bte) * live, acc_p);
                                const float u3 = g4 * x3;            // g4 * x3
                                srow += u3;
                                acc_c = fmaf(u3, q1, acc_c);
                                const float g3x3 = u3 * m;           // g3 * x3
                                acc_g = fmaf(g3x3, l2, acc_g);       // * ln2 applied in finalize
                                const float g2 = x2 >= kGammaClamp ? g3x3 * pg * rcp_fast(x2c) : 0.f;
                                acc_s = fmaf(g2, x1, acc_s);
                                const float g1 = g2 * ps;
                                // d x1 / d w = (x0 - a) * ica / txc^2   (tx = 1 - w ica, only where tx >= 0.01)
                                if (HAS_ICA) {
                                    if (pass_tx) acc_w = fmaf(g1 * xa, icae * inv * inv, acc_w);
                                } else {
                                    acc_w = fmaf(g1, xa, acc_w);     // * ica * inv^2 applied in finalize
                                }
                                dxv[e] = g1 * inv;
                            }
                            if (dx) {
                                float* d = dx + (size_t)u.plane * H * W + off;
                                d[0] = dxv[0];
                                if (two) d[1] = dxv[1];
                            }
                        }
                        SS[(8 * rg + r) * kSP + cp] = srow;
                    }
                }
            }
            __syncthreads();
            {   // row sums S: 32 rows x 64 partials -> 8 lanes per row, fixed order
                const int rr = tid >> 3, part8 = tid & 7;
                const float4 s0 = *reinterpret_cast<const float4*>(SS + rr * kSP + 8 * part8);
                const float4 s1 = *reinterpret_cast<const float4*>(SS + rr * kSP + 8 * part8 + 4);
                float s = ((s0.x + s0.y) + (s0.z + s0.w)) + ((s1.x + s1.y) + (s1.z + s1.w));
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                const int o = n * kRB - kRadius + rr;
                if (part8 == 0 && o >= kRadius && o < kRadius + u.seg_len)
                    Spart[((size_t)u.plane * H + u.r0 + o - kRadius) * sc.strips + u.strip] = s;
            }
        }
    }
    if (cur_ps >= 0) flush();
}

// One CTA per image: fixed-order sum of the (CTA, plane-strip) partials, the row-coupled fix-up of columns 0..2
// (d lum / d x3[:, :, :, 0..2]) and the regressor Jacobians -> dfeat[b, 0..14].
template <bool HAS_ICA, bool FAST>
__global__ void __launch_bounds__(kThreads)
recovery_bwd_finalize_kernel(const float* __restrict__ x, const float* __restrict__ A,
                             const float* __restrict__ IcA, const float* __restrict__ feat,
                             const float* __restrict__ part, const float* __restrict__ Spart,
                             float* __restrict__ dfeat, float* __restrict__ dx, int B, int H, int W) {
    __shared__ ImgParams sp;
    __shared__ double s_red[32];
    const int tid = threadIdx.x, b = blockIdx.x;
    const Sched sc = make_sched(B, H, W);
    if (tid == 0) regress(feat + b * kFeat, sp);
    __syncthreads();
    const float pg = sp.gamma, pc = sp.c;
    double dp = 0, dc = 0, dg = 0, dw = 0, ds[3] = {0, 0, 0};

    // partial sums: plane-strip ps of this image was processed by CTAs c_of(first block) .. c_of(last block)
    for (int i = tid; i < 3 * sc.strips; i += kThreads) {
        const int ps = 3 * b * sc.strips + i;
        const int ch = i / sc.strips;
        const long long x0 = (long long)ps * sc.nRB, x1 = x0 + sc.nRB - 1;
        const int c_first = (int)(((x0 + 1) * sc.G + sc.N - 1) / sc.N) - 1;
        const int c_last = (int)(((x1 + 1) * sc.G + sc.N - 1) / sc.N) - 1;
        // default-IcA constants folded out of the kernel's acc_w
        const float txc = fmaxf(1.f - sp.w * kDefaultIcA, kTxMin);
        const float wk = HAS_ICA ? 1.f : ((1.f - sp.w * kDefaultIcA >= kTxMin) ? kDefaultIcA / (txc * txc) : 0.f);
        for (int c = c_first; c <= c_last; ++c) {
            const float* q = part + (size_t)(c + ps) * kBwdSums;
            dp += q[0]; dc += q[1]; dg += (double)q[2] * 0.69314718055994530942; ds[ch] += q[3]; dw += (double)q[4] * wk;
        }
    }
    const float kappa[3] = {kLumR, kLumG, kLumB};
    for (int i = tid; i < 3 * H; i += kThreads) {
        const int ch = i / H, row = i - ch * H, plane = 3 * b + ch;
        float S = 0.f;
        for (int st = 0; st < sc.strips; ++st) S += Spart[((size_t)plane * H + row) * sc.strips + st];
        const float a = A ? __ldg(A + b * 3 + ch) : kDefaultA;
        const ChainK ck = make_chain(sp, ch, a);
        const size_t off = ((size_t)plane * H + row) * W;
        float x0[3], ica[3], tx[3], txc[3], x1[3], x2[3], x2c[3], x3[3], l2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x0[k] = __ldg(x + off + k);
            ica[k] = HAS_ICA ? __ldg(IcA + ((size_t)b * H + row) * W + k) : kDefaultIcA;
            tx[k] = 1.f - ck.w * ica[k];
            txc[k] = fmaxf(tx[k], kTxMin);
            x1[k] = (x0[k] - a) / txc[k] + a;
            x2[k] = x1[k] * ck.s;
            x2c[k] = fmaxf(x2[k], kGammaClamp);
            x3[k] = gamma_pow<FAST>(x2c[k], pg, &l2[k]);
        }
        const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
        if (rl.lraw >= 0.f && rl.lraw <= 1.f) {
            const float dq = 0.5f * kPi * sinf(kPi * rl.lum) / rl.denom - rl.cl / (rl.denom * rl.denom);
            const float glum = pc * dq * S;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float e3 = kappa[k] * glum;
                dg += (double)(e3 * x3[k] * l2[k]) * 0.69314718055994530942;
                const float e2 = x2[k] >= kGammaClamp ? e3 * pg * x3[k] / x2c[k] : 0.f;
                ds[ch] += (double)(e2 * x1[k]);
                const float e1 = e2 * ck.s;
                if (tx[k] >= kTxMin) dw += (double)(e1 * (x0[k] - a) * ica[k] / (txc[k] * txc[k]));
                if (dx) dx[off + k] += e1 / txc[k];
            }
        }
    }
    dp = block_sum<double>(dp, s_red);
    dc = block_sum<double>(dc, s_red);
    dg = block_sum<double>(dg, s_red);
    dw = block_sum<double>(dw, s_red);
    ds[0] = block_sum<double>(ds[0], s_red);
    ds[1] = block_sum<double>(ds[1], s_red);
    ds[2] = block_sum<double>(ds[2], s_red);
    if (tid == 0) {
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
}

constexpr size_t kFwdSmem = (size_t)(kRing * kXP + kRing * kHP + kMaxU) * sizeof(float);
constexpr size_t kBwdSmem = (size_t)(kRing * kXP + kRing * kHP + 2 * kMaxU + kRB * kSP) * sizeof(float);

static int check_shape(const char* who, int B, int H, int W) {
    DD_REQUIRE(B > 0 && H > 0 && W > 0, DD_ERR_INVALID, "%s: B, H, W must be positive (got %d, %d, %d)", who, B, H, W);
    DD_REQUIRE(W >= 3, DD_ERR_WIDTH_LT3, "%s: W = %d < 3: rgb2lum indexes columns 0..2", who, W);
    DD_REQUIRE(H > kRadius && W > kRadius, DD_ERR_REFLECT_PAD,
               "%s: reflect padding of 12 needs H, W > 12 (got %d x %d)", who, H, W);
    return DD_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return DD_ERR_CUDA;
    }
    return DD_OK;
}

static bool precise_mode() {  // DEDARK_PRECISE=1: powf/logf instead of the MUFU pow (debugging aid, read once)
    static const bool v = [] { const char* e = getenv("DEDARK_PRECISE"); return e && e[0] == '1'; }();
    return v;
}

template <bool HAS_ICA, bool FAST>
static int launch_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H, int W,
                      cudaStream_t st) {
    const Sched sc = make_sched(B, H, W);
    if (int e = set_smem(recovery_fwd_kernel<HAS_ICA, FAST>, kFwdSmem)) return e;
    recovery_fwd_kernel<HAS_ICA, FAST><<<sc.G, kThreads, kFwdSmem, st>>>(x, A, IcA, feat, y, B, H, W);
    count_launch();
    return check_launch("dd_recovery_fwd");
}

template <bool HAS_ICA, bool FAST>
static int launch_bwd(const float* x, const float* A, const float* IcA, const float* feat, const float* g, float* dfeat,
                      float* dx, int B, int H, int W, float* ws, cudaStream_t st) {
    const Sched sc = make_sched(B, H, W);
    float* part = ws;
    float* Spart = part + (size_t)(sc.G + sc.nPS) * kBwdSums;
    if (int e = set_smem(recovery_bwd_kernel<HAS_ICA, FAST>, kBwdSmem)) return e;
    recovery_bwd_kernel<HAS_ICA, FAST><<<sc.G, kThreads, kBwdSmem, st>>>(x, A, IcA, feat, g, part, Spart, dx, B, H, W);
    recovery_bwd_finalize_kernel<HAS_ICA, FAST><<<B, kThreads, 0, st>>>(x, A, IcA, feat, part, Spart, dfeat, dx, B, H, W);
    count_launch(2);
    return check_launch("dd_recovery_bwd");
}

}  // namespace dd

extern "C" int dd_recovery_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B,
                               int H, int W, void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_shape("dd_recovery_fwd", B, H, W)) return e;
    DD_REQUIRE(x && feat && y && x != y, DD_ERR_INVALID, "dd_recovery_fwd: null pointer or y aliases x");
    const bool fast = !precise_mode();
    if (IcA) return fast ? launch_fwd<true, true>(x, A, IcA, feat, y, B, H, W, st) : launch_fwd<true, false>(x, A, IcA, feat, y, B, H, W, st);
    return fast ? launch_fwd<false, true>(x, A, nullptr, feat, y, B, H, W, st) : launch_fwd<false, false>(x, A, nullptr, feat, y, B, H, W, st);
}

extern "C" int dd_recovery_bwd(const float* x, const float* A, const float* IcA, const float* feat, const float* g_,
                               float* dfeat, float* dx, int B, int H, int W, void* ws, size_t ws_bytes,
                               void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_shape("dd_recovery_bwd", B, H, W)) return e;
    DD_REQUIRE(x && feat && g_ && dfeat, DD_ERR_INVALID, "dd_recovery_bwd: null pointer");
    DD_REQUIRE(ws && ws_bytes >= recovery_bwd_ws_bytes(B, H, W), DD_ERR_WORKSPACE,
               "dd_recovery_bwd: workspace %zu < %zu", ws_bytes, recovery_bwd_ws_bytes(B, H, W));
    const bool fast = !precise_mode();
    float* w = reinterpret_cast<float*>(ws);
    if (IcA)
        return fast ? launch_bwd<true, true>(x, A, IcA, feat, g_, dfeat, dx, B, H, W, w, st)
                    : launch_bwd<true, false>(x, A, IcA, feat, g_, dfeat, dx, B, H, W, w, st);
    return fast ? launch_bwd<false, true>(x, A, nullptr, feat, g_, dfeat, dx, B, H, W, w, st)
                : launch_bwd<false, false>(x, A, nullptr, feat, g_, dfeat, dx, B, H, W, w, st);
}
