// dd_recovery.cu -- a6..a12 + a14: the fused filter chain (regressors, DeDark -> WB -> Gamma ->
// Contrast -> USM), forward and backward, one pass over the image each.
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,
// 295-304,316-317.  Closed-form backward: SURVEY.md section 8(a) row a14 (checked against autograd).
//
// Work decomposition ("marching strips").  A unit is (image plane b*3+ch, 128-column strip, row
// segment).  One CTA of 256 threads marches down its strip in blocks of 16 rows:
//     1. pointwise chain x0 -> x4 for 16 rows x (128+24) columns        registers -> smem ring XS
//     2. horizontal 25-tap pass, 8 outputs per thread from 32 staged taps   XS -> smem ring HS
//     3. vertical 25-tap pass, 8 rows per thread, then y = (x4 - blur) p + x4 -> global
// so the 12-row vertical halo is paid once per segment (not per tile) and the pointwise chain (one
// powf per pixel) is recomputed only on the 24 halo columns of a strip.  Row scalars of the
// row-coupled contrast filter (the rgb2lum quirk: "luminance" = columns 0..2 of each row) are
// computed once per CTA.  The rings hold 48 rows: XS pitch 156 and HS pitch 132 floats make the
// 128-bit shared-memory accesses of pass 2 bank-conflict-free (lanes map to rows).
//
// Backward uses the same skeleton on g = dL/dy: the ring holds g (zero outside the image), the two
// passes apply the ADJOINT of the reflect-padded blur (zero-padded correlation + the reflect
// fold-back as a gather: W(i,j) = k[j-i] + [j>=1] k[j+i] + [j<=n-2] k[2(n-1)-j-i]); the chain is
// recomputed from x and the seven parameter gradients are reduced per thread -> per CTA partials ->
// a fixed-order finalize kernel (no float atomics), which also applies the row-coupled fix-up of
// columns 0..2 and the regressor Jacobians.
#include "dd_common.cuh"
#include "dd_layout.cuh"

namespace dd {

constexpr int kThreads = 256;
constexpr int kRB = 16;                       // rows per marching block
constexpr int kRing = 48;                     // ring depth (rows)
constexpr int kXW = kStripW + 2 * kRadius;    // 152 staged columns
constexpr int kXP = 156;                      // XS pitch (floats): (kXP/4) odd -> conflict-free LDS.128 by row
constexpr int kHP = 132;                      // HS pitch (floats)
constexpr int kPre = (kRB * kXW + kThreads - 1) / kThreads;  // staged elements per thread per block (10)
constexpr int kMaxU = kMaxSegRows + 2 * kRadius;

__constant__ float c_tap[13] = {DD_K0, DD_K1, DD_K2, DD_K3, DD_K4, DD_K5, DD_K6,
                                DD_K7, DD_K8, DD_K9, DD_K10, DD_K11, DD_K12};

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

struct Chain {
    float tx, txc, x1, x2c, x3;
    bool pass_gamma;  // x2 >= 1e-4 (gradient passes the clamp)
};

// DeDark (filtersB.py:211-214) -> WB (:259) -> Gamma (:232-233) for one pixel
__device__ __forceinline__ Chain chain(float x0, float a, float ica, float w, float s, float gamma) {
    Chain r;
    r.tx = 1.f - w * ica;
    r.txc = fmaxf(r.tx, kTxMin);
    r.x1 = (x0 - a) / r.txc + a;
    const float x2 = r.x1 * s;
    r.pass_gamma = x2 >= kGammaClamp;
    r.x2c = fmaxf(x2, kGammaClamp);
    r.x3 = powf(r.x2c, gamma);
    return r;
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

// 8 horizontally adjacent outputs from 32 staged values (window of output t = in[t .. t+24])
__device__ __forceinline__ void hpass8(const float* __restrict__ xrow, float out[8]) {
    float in[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(xrow + 4 * i);
        in[4 * i] = v.x; in[4 * i + 1] = v.y; in[4 * i + 2] = v.z; in[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) a = fmaf(in[t + j], tap(j < kRadius ? kRadius - j : j - kRadius), a);
        out[t] = a;
    }
}

// 8 vertically adjacent outputs at one column; hv[i] = ring row (first - 12 + i)
__device__ __forceinline__ void vpass8(const float* __restrict__ HS, int base_slot, int col, float out[8]) {
    float hv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        int s = base_slot + i;
        s -= s >= kRing ? kRing : 0;
        hv[i] = HS[s * kHP + col];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) a = fmaf(hv[r + j], tap(j < kRadius ? kRadius - j : j - kRadius), a);
        out[r] = a;
    }
}

struct Unit {
    int plane, b, ch, strip, seg, r0, r1, c0, nU, nB;
};
__device__ __forceinline__ Unit decode_unit(int unit, int H, int strips, int segs, int seg_rows) {
    Unit u;
    u.strip = unit % strips;
    u.seg = (unit / strips) % segs;
    u.plane = unit / (strips * segs);
    u.b = u.plane / 3;
    u.ch = u.plane % 3;
    u.r0 = u.seg * seg_rows;
    u.r1 = min(H, u.r0 + seg_rows);
    u.c0 = u.strip * kStripW;
    u.nU = u.r1 - u.r0 + 2 * kRadius;
    u.nB = (u.nU + kRB - 1) / kRB;
    return u;
}

// =================================================================================================
// forward
// =================================================================================================
template <bool HAS_ICA>
__global__ void __launch_bounds__(kThreads)
recovery_fwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                    const float* __restrict__ feat, float* __restrict__ y, int H, int W, int strips, int segs,
                    int seg_rows) {
    extern __shared__ __align__(16) float smem[];
    float* XS = smem;
    float* HS = XS + kRing * kXP;
    float* MS = HS + kRing * kHP;  // per virtual row: m = (1-c) + c*q
    __shared__ ImgParams sp;

    const int tid = threadIdx.x;
    const Unit u = decode_unit(blockIdx.x, H, strips, segs, seg_rows);
    if (tid == 0) regress(feat + u.b * kFeat, sp);
    __syncthreads();
    const float pw = sp.w, ps = sp.s[u.ch], pg = sp.gamma, pc = sp.c, pp = sp.p;
    const float a = A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA;
    const float* xp = x + (size_t)u.plane * H * W;
    const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;

    for (int v = tid; v < u.nU; v += kThreads) {
        const int row = reflect(u.r0 - kRadius + v, H);
        float x3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ica = HAS_ICA ? __ldg(ip + (size_t)row * W + k) : kDefaultIcA;
            x3[k] = chain(__ldg(xp + (size_t)row * W + k), a, ica, pw, ps, pg).x3;
        }
        const RowLum rl = row_lum<true>(x3[0], x3[1], x3[2]);
        MS[v] = (1.f - pc) + pc * rl.q;
    }

    float pre[kPre], prei[kPre];
    auto stage = [&](int n) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int idx = tid + k * kThreads;
            const int rr = idx / kXW, cc = idx - rr * kXW;
            const int v = n * kRB + rr;
            if (idx < kRB * kXW && v < u.nU) {
                const int row = reflect(u.r0 - kRadius + v, H);
                int gc = reflect(u.c0 - kRadius + cc, W);
                gc = min(max(gc, 0), W - 1);
                pre[k] = __ldg(xp + (size_t)row * W + gc);
                if (HAS_ICA) prei[k] = __ldg(ip + (size_t)row * W + gc);
            }
        }
    };
    stage(0);

    const int seg_len = u.r1 - u.r0;
    for (int n = 0; n < u.nB; ++n) {
        __syncthreads();  // MS ready (n == 0); ring slots of block n are no longer read by the previous V pass
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int idx = tid + k * kThreads;
            const int rr = idx / kXW, cc = idx - rr * kXW;
            const int v = n * kRB + rr;
            if (idx < kRB * kXW && v < u.nU) {
                const float ica = HAS_ICA ? prei[k] : kDefaultIcA;
                const float x3 = chain(pre[k], a, ica, pw, ps, pg).x3;
                XS[(v % kRing) * kXP + cc] = x3 * MS[v];
            }
        }
        if (n + 1 < u.nB) stage(n + 1);
        __syncthreads();
        {   // horizontal pass: lanes -> rows
            const int rr = tid & 15, cg = tid >> 4;
            const int slot = (n * kRB + rr) % kRing;
            float o[8];
            hpass8(XS + slot * kXP + 8 * cg, o);
            float4* dst = reinterpret_cast<float4*>(HS + slot * kHP + 8 * cg);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        {   // vertical pass + USM epilogue: lanes -> columns
            const int col = tid & 127, half = tid >> 7;
            const int o_first = n * kRB - kRadius + 8 * half;  // virtual row of the first output
            if (o_first >= kRadius && o_first < kRadius + seg_len) {
                float bl[8];
                vpass8(HS, (o_first - kRadius) % kRing, col, bl);
                if (u.c0 + col < W) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int o = o_first + r;
                        if (o < kRadius + seg_len) {
                            const float x4 = XS[(o % kRing) * kXP + col + kRadius];
                            y[((size_t)u.plane * H + u.r0 + o - kRadius) * W + u.c0 + col] = (x4 - bl[r]) * pp + x4;
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
template <bool HAS_ICA>
__global__ void __launch_bounds__(kThreads)
recovery_bwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                    const float* __restrict__ feat, const float* __restrict__ g, float* __restrict__ part,
                    float* __restrict__ Spart, float* __restrict__ dx, int H, int W, int strips, int segs,
                    int seg_rows) {
    extern __shared__ __align__(16) float smem[];
    float* XS = smem;                 // g, zero outside the image
    float* HS = XS + kRing * kXP;
    float* MSm = HS + kRing * kHP;    // per virtual row: m = (1-c) + c*q
    float* MSq = MSm + kMaxU;         // per virtual row: q = cl / (lum + eps)
    __shared__ ImgParams sp;
    __shared__ float sS[kRB][4];
    __shared__ float s_red[32];

    const int tid = threadIdx.x;
    const Unit u = decode_unit(blockIdx.x, H, strips, segs, seg_rows);
    if (tid == 0) regress(feat + u.b * kFeat, sp);
    __syncthreads();
    const float pw = sp.w, ps = sp.s[u.ch], pg = sp.gamma, pc = sp.c, pp = sp.p;
    const float a = A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA;
    const float* xp = x + (size_t)u.plane * H * W;
    const float* gp = g + (size_t)u.plane * H * W;
    const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
    const int seg_len = u.r1 - u.r0;

    for (int v = tid; v < u.nU; v += kThreads) {
        const int row = u.r0 - kRadius + v;
        float m = 0.f, q = 0.f;
        if (row >= 0 && row < H) {
            float x3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float ica = HAS_ICA ? __ldg(ip + (size_t)row * W + k) : kDefaultIcA;
                x3[k] = chain(__ldg(xp + (size_t)row * W + k), a, ica, pw, ps, pg).x3;
            }
            const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
            q = rl.q;
            m = (1.f - pc) + pc * rl.q;
        }
        MSm[v] = m;
        MSq[v] = q;
    }

    float pre[kPre];
    auto stage = [&](int n) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int idx = tid + k * kThreads;
            const int rr = idx / kXW, cc = idx - rr * kXW;
            const int row = u.r0 - kRadius + n * kRB + rr;
            const int gc = u.c0 - kRadius + cc;
            pre[k] = (idx < kRB * kXW && row >= 0 && row < H && gc >= 0 && gc < W) ? __ldg(gp + (size_t)row * W + gc) : 0.f;
        }
    };
    stage(0);

    float acc_p = 0.f, acc_c = 0.f, acc_g = 0.f, acc_s = 0.f, acc_w = 0.f;

    for (int n = 0; n < u.nB; ++n) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int idx = tid + k * kThreads;
            const int rr = idx / kXW, cc = idx - rr * kXW;
            if (idx < kRB * kXW) XS[((n * kRB + rr) % kRing) * kXP + cc] = pre[k];
        }
        if (n + 1 < u.nB) stage(n + 1);
        __syncthreads();
        {   // horizontal adjoint pass
            const int rr = tid & 15, cg = tid >> 4;
            const int slot = (n * kRB + rr) % kRing;
            const float* xrow = XS + slot * kXP;
            float o[8];
            hpass8(xrow + 8 * cg, o);
            const int j0 = u.c0 + 8 * cg;  // global column of o[0]
            if (j0 <= kRadius || j0 + 7 >= W - 1 - kRadius) {  // reflect fold-back (border strips only)
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
            float4* dst = reinterpret_cast<float4*>(HS + slot * kHP + 8 * cg);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        __syncthreads();
        {   // vertical adjoint pass + chain recompute + reductions
            const int col = tid & 127, half = tid >> 7, lane = tid & 31, wq = (tid >> 5) & 3;
            const int o_first = n * kRB - kRadius + 8 * half;
            const bool group_ok = o_first >= kRadius && o_first < kRadius + seg_len;
            float bt[8];
            if (group_ok) vpass8(HS, (o_first - kRadius) % kRing, col, bt);
            const bool col_ok = u.c0 + col < W;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int o = o_first + r;
                const int jr = u.r0 + o - kRadius;  // image row
                float sval = 0.f;
                if (group_ok && col_ok && o < kRadius + seg_len) {
                    float b = bt[r];
                    if (jr >= 1 && jr <= kRadius)
                        for (int i = 0; i <= kRadius - jr; ++i)
                            b = fmaf(HS[((i - u.r0 + kRadius) % kRing) * kHP + col], c_tap[jr + i], b);
                    if (jr >= H - 1 - kRadius && jr <= H - 2)
                        for (int i = max(0, 2 * (H - 1) - jr - kRadius); i <= H - 1; ++i)
                            b = fmaf(HS[((i - u.r0 + kRadius) % kRing) * kHP + col], c_tap[abs(2 * (H - 1) - jr - i)], b);
                    const float g5 = XS[(o % kRing) * kXP + col + kRadius];
                    const size_t off = (size_t)jr * W + u.c0 + col;
                    const float x0 = __ldg(xp + off);
                    const float ica = HAS_ICA ? __ldg(ip + off) : kDefaultIcA;
                    const Chain c = chain(x0, a, ica, pw, ps, pg);
                    const float m = MSm[o], q = MSq[o];
                    const float x4 = c.x3 * m;
                    const float g4 = g5 * (1.f + pp) - pp * b;
                    acc_p = fmaf(x4, g5 - b, acc_p);
                    acc_c = fmaf(g4 * c.x3, q - 1.f, acc_c);
                    sval = g4 * c.x3;
                    const float g3 = g4 * m;
                    acc_g = fmaf(g3 * c.x3, logf(c.x2c), acc_g);
                    const float g2 = c.pass_gamma ? g3 * pg * c.x3 / c.x2c : 0.f;
                    acc_s = fmaf(g2, c.x1, acc_s);
                    const float g1 = g2 * ps;
                    if (c.tx >= kTxMin) acc_w = fmaf(g1 * (x0 - a), ica / (c.txc * c.txc), acc_w);
                    if (dx) dx[(size_t)u.plane * H * W + off] = g1 / c.txc;
                }
                sval = warp_sum(sval);
                if (lane == 0) sS[8 * half + r][wq] = sval;
            }
        }
        __syncthreads();
        if (tid < kRB) {
            const int o = n * kRB - kRadius + tid;
            if (o >= kRadius && o < kRadius + seg_len)
                Spart[((size_t)u.plane * H + u.r0 + o - kRadius) * strips + u.strip] =
                    (sS[tid][0] + sS[tid][1]) + (sS[tid][2] + sS[tid][3]);
        }
    }
    float* out = part + (size_t)blockIdx.x * kBwdSums;
    float s;
    s = block_sum<float>(acc_p, s_red); if (tid == 0) out[0] = s;
    s = block_sum<float>(acc_c, s_red); if (tid == 0) out[1] = s;
    s = block_sum<float>(acc_g, s_red); if (tid == 0) out[2] = s;
    s = block_sum<float>(acc_s, s_red); if (tid == 0) out[3] = s;
    s = block_sum<float>(acc_w, s_red); if (tid == 0) out[4] = s;
}

// One CTA per image: fixed-order sum of the unit partials, the row-coupled fix-up of columns 0..2
// (d lum / d x3[:, :, :, 0..2]) and the regressor Jacobians -> dfeat[b, 0..14].
template <bool HAS_ICA>
__global__ void __launch_bounds__(kThreads)
recovery_bwd_finalize_kernel(const float* __restrict__ x, const float* __restrict__ A,
                             const float* __restrict__ IcA, const float* __restrict__ feat,
                             const float* __restrict__ part, const float* __restrict__ Spart,
                             float* __restrict__ dfeat, float* __restrict__ dx, int H, int W, int strips, int segs) {
    __shared__ ImgParams sp;
    __shared__ double s_red[32];
    const int tid = threadIdx.x, b = blockIdx.x;
    if (tid == 0) regress(feat + b * kFeat, sp);
    __syncthreads();
    const float pw = sp.w, pg = sp.gamma, pc = sp.c;
    double dp = 0, dc = 0, dg = 0, dw = 0, ds[3] = {0, 0, 0};

    const int upp = segs * strips;  // units per plane
    for (int i = tid; i < 3 * upp; i += kThreads) {
        const float* q = part + ((size_t)3 * b * upp + i) * kBwdSums;
        dp += q[0]; dc += q[1]; dg += q[2]; ds[i / upp] += q[3]; dw += q[4];
    }
    const float kappa[3] = {kLumR, kLumG, kLumB};
    for (int i = tid; i < 3 * H; i += kThreads) {
        const int ch = i / H, row = i - ch * H, plane = 3 * b + ch;
        float S = 0.f;
        for (int st = 0; st < strips; ++st) S += Spart[((size_t)plane * H + row) * strips + st];
        const float a = A ? __ldg(A + b * 3 + ch) : kDefaultA;
        const float ps = sp.s[ch];
        const size_t off = ((size_t)plane * H + row) * W;
        Chain c[3];
        float x0[3], ica[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x0[k] = __ldg(x + off + k);
            ica[k] = HAS_ICA ? __ldg(IcA + ((size_t)b * H + row) * W + k) : kDefaultIcA;
            c[k] = chain(x0[k], a, ica[k], pw, ps, pg);
        }
        const RowLum rl = row_lum<false>(c[0].x3, c[1].x3, c[2].x3);
        if (rl.lraw >= 0.f && rl.lraw <= 1.f) {
            const float dq = 0.5f * kPi * sinf(kPi * rl.lum) / rl.denom - rl.cl / (rl.denom * rl.denom);
            const float glum = pc * dq * S;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float e3 = kappa[k] * glum;
                dg += (double)(e3 * c[k].x3 * logf(c[k].x2c));
                const float e2 = c[k].pass_gamma ? e3 * pg * c[k].x3 / c[k].x2c : 0.f;
                ds[ch] += (double)(e2 * c[k].x1);
                const float e1 = e2 * ps;
                if (c[k].tx >= kTxMin) dw += (double)(e1 * (x0[k] - a) * ica[k] / (c[k].txc * c[k].txc));
                if (dx) dx[off + k] += e1 / c[k].txc;
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
constexpr size_t kBwdSmem = (size_t)(kRing * kXP + kRing * kHP + 2 * kMaxU) * sizeof(float);

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

}  // namespace dd

extern "C" int dd_recovery_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B,
                               int H, int W, void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_shape("dd_recovery_fwd", B, H, W)) return e;
    DD_REQUIRE(x && feat && y && x != y, DD_ERR_INVALID, "dd_recovery_fwd: null pointer or y aliases x");
    const RecoveryGrid g = recovery_grid(B, H, W);
    if (IcA) {
        if (int e = set_smem(recovery_fwd_kernel<true>, kFwdSmem)) return e;
        recovery_fwd_kernel<true><<<g.units, kThreads, kFwdSmem, st>>>(x, A, IcA, feat, y, H, W, g.strips, g.segs, g.seg_rows);
    } else {
        if (int e = set_smem(recovery_fwd_kernel<false>, kFwdSmem)) return e;
        recovery_fwd_kernel<false><<<g.units, kThreads, kFwdSmem, st>>>(x, A, nullptr, feat, y, H, W, g.strips, g.segs, g.seg_rows);
    }
    count_launch();
    return check_launch("dd_recovery_fwd");
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
    const RecoveryGrid g = recovery_grid(B, H, W);
    float* part = reinterpret_cast<float*>(ws);
    float* Spart = part + (size_t)g.units * kBwdSums;
    if (IcA) {
        if (int e = set_smem(recovery_bwd_kernel<true>, kBwdSmem)) return e;
        recovery_bwd_kernel<true><<<g.units, kThreads, kBwdSmem, st>>>(x, A, IcA, feat, g_, part, Spart, dx, H, W, g.strips, g.segs, g.seg_rows);
        recovery_bwd_finalize_kernel<true><<<B, kThreads, 0, st>>>(x, A, IcA, feat, part, Spart, dfeat, dx, H, W, g.strips, g.segs);
    } else {
        if (int e = set_smem(recovery_bwd_kernel<false>, kBwdSmem)) return e;
        recovery_bwd_kernel<false><<<g.units, kThreads, kBwdSmem, st>>>(x, A, nullptr, feat, g_, part, Spart, dx, H, W, g.strips, g.segs, g.seg_rows);
        recovery_bwd_finalize_kernel<false><<<B, kThreads, 0, st>>>(x, A, nullptr, feat, part, Spart, dfeat, dx, H, W, g.strips, g.segs);
    }
    count_launch(2);
    return check_launch("dd_recovery_bwd");
}
