// dd_recovery_fwd.cu -- a6..a12: the fused filter chain forward (regressors, DeDark -> WB -> Gamma -> Contrast ->
// USM) in one pass over the image.
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,295-304,316-317.
//
// "Marching strips" (dd_layout.cuh): every (plane, 128-column strip) is cut into row-blocks of 32 rows; 296
// persistent CTAs (2 per SM, 256 threads) each own a contiguous, equal range of row-blocks and march down them:
//   1. stage   x0 -> pointwise chain -> x4 for 32 rows x (128+24) columns.  Global loads are 128-bit and issued one
//              block ahead (register prefetch); columns outside the image are not loaded at all -- the reflect
//              halo is filled by mirroring the freshly computed x4 values inside shared memory.      -> ring XS
//   2. H pass  25 taps, 8 outputs per thread from 8 LDS.128, FFMA2 with immediate taps.              -> ring HS
//   3. V pass  25 taps, 2 columns x 8 rows per thread from 32 LDS.64, FFMA2, then y = (x4 - blur) p + x4 written
//              with 64-bit coalesced stores.
// The vertical halo (24 rows) is paid once per CTA range.  XS pitch 156 / HS pitch 132 floats make the 128-bit
// shared accesses of pass 2 conflict-free (lanes map to rows).  The blur is FMA-pipe bound (50 FMA per pixel ~ 28 us
// per 16x3x640^2 at the measured 35 TFMA/s); FFMA2 halves its issue slots so loads, MUFU and index arithmetic issue
// in the shadow of the FMA pipe.
#include "dd_recovery.cuh"

namespace dd {

constexpr int kXRingF = 64;  // XS ring depth of the forward kernel (rows), power of two

template <bool HAS_ICA, bool FAST, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 2)
recovery_fwd_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                    const float* __restrict__ feat, float* __restrict__ y, int B, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* XS = smem;
    float* HS = XS + kXRingF * kXP;
    float* MS = HS + kHRing * kHP;  // per virtual row: m = (1-c) + c*q
    __shared__ ImgParams sp;

    const int tid = threadIdx.x;
    const Sched sc = make_sched(B, H, W);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const bool w2 = (W & 1) == 0;
    const int L = W - 1;

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
                const int gc = u.c0 - kRadius + 4 * c4;
                if (f < kRB * kXW4 && v < u.nU) {
                    const size_t ro = (size_t)reflect(u.r0 - kRadius + v, H) * W;
                    if (ALIGNED) {
                        if (gc >= 0 && gc < W) {
                            pre[k] = __ldg(reinterpret_cast<const float4*>(xp + ro + gc));
                            if (HAS_ICA) prei[k] = __ldg(reinterpret_cast<const float4*>(ip + ro + gc));
                        }
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
                const int gc = u.c0 - kRadius + 4 * c4;
                if (f < kRB * kXW4 && v < u.nU && (!ALIGNED || (gc >= 0 && gc < W))) {
                    const float m = MS[v];
                    const float4 in = pre[k];
                    const float4 ic = HAS_ICA ? prei[k] : make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                    float o[4];
                    o[0] = chain_x3<HAS_ICA, FAST>(ck, in.x, ic.x) * m;
                    o[1] = chain_x3<HAS_ICA, FAST>(ck, in.y, ic.y) * m;
                    o[2] = chain_x3<HAS_ICA, FAST>(ck, in.z, ic.z) * m;
                    o[3] = chain_x3<HAS_ICA, FAST>(ck, in.w, ic.w) * m;
                    float* xrow = XS + (v & (kXRingF - 1)) * kXP;
                    *reinterpret_cast<float4*>(xrow + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
                    if (ALIGNED) {
                        // reflect halo (F.pad mode='reflect', filtersB.py:167): image col -j <- j, L+d <- L-d
                        if (gc <= kRadius) {  // only strip 0 has image columns 0..12 at gc <= 12
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = gc + e;
                                if (j >= 1 && j <= kRadius) xrow[kRadius - j - u.c0] = o[e];
                            }
                        }
                        if (gc + 3 >= L - kRadius) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int d = L - (gc + e);
                                const int tc = L + d - u.c0 + kRadius;
                                if (d >= 1 && d <= kRadius && tc < kXW) xrow[tc] = o[e];
                            }
                        }
                    }
                }
            }
            if (n + 1 < u.nB) stage(n + 1);
            __syncthreads();
            {   // horizontal pass: lanes -> rows, each thread two groups of 8 columns
                const int rr = tid & 31, cg = tid >> 5;
                const int slot = (n * kRB + rr) & (kXRingF - 1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cb = 64 * half + 8 * cg;
                    float o[8];
                    hpass8(XS + slot * kXP + cb, o);
                    float4* dst = reinterpret_cast<float4*>(HS + ((n * kRB + rr) & (kHRing - 1)) * kHP + cb);
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
                    vpass8x2(HS, (o_first - kRadius) & (kHRing - 1), col2, bl);
                    const int gc = u.c0 + col2;
                    const u64 p2 = pk(pp, pp), m1 = pk(-1.f, -1.f);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int o = o_first + r;
                        if (o < kRadius + u.seg_len && gc < W) {
                            const u64 x4 = *reinterpret_cast<const u64*>(XS + (o & (kXRingF - 1)) * kXP + col2 + kRadius);
                            const u64 yv = fma2(fma2(bl[r], m1, x4), p2, x4);  // (x4 - blur) * p + x4
                            float* dst = yp + (size_t)(u.r0 + o - kRadius) * W + gc;
                            if (w2) {
                                __stcs(reinterpret_cast<float2*>(dst), upk(yv));
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

constexpr size_t kFwdSmem = (size_t)(kXRingF * kXP + kHRing * kHP + kMaxU) * sizeof(float);

template <bool HAS_ICA, bool FAST, bool ALIGNED>
static int launch_fwd3(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H,
                       int W, cudaStream_t st) {
    const Sched sc = make_sched(B, H, W);
    if (int e = set_smem(recovery_fwd_kernel<HAS_ICA, FAST, ALIGNED>, kFwdSmem)) return e;
    recovery_fwd_kernel<HAS_ICA, FAST, ALIGNED><<<sc.G, kThreads, kFwdSmem, st>>>(x, A, IcA, feat, y, B, H, W);
    count_launch();
    return check_launch("dd_recovery_fwd");
}

template <bool HAS_ICA, bool FAST>
static int launch_fwd2(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H,
                       int W, cudaStream_t st) {
    const bool aligned = (W & 3) == 0 && ((uintptr_t)x & 15) == 0 && (!HAS_ICA || ((uintptr_t)IcA & 15) == 0);
    return aligned ? launch_fwd3<HAS_ICA, FAST, true>(x, A, IcA, feat, y, B, H, W, st)
                   : launch_fwd3<HAS_ICA, FAST, false>(x, A, IcA, feat, y, B, H, W, st);
}

}  // namespace dd

extern "C" int dd_recovery_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B,
                               int H, int W, void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_fwd", B, H, W)) return e;
    DD_REQUIRE(x && feat && y && x != y, DD_ERR_INVALID, "dd_recovery_fwd: null pointer or y aliases x");
    DD_REQUIRE(((uintptr_t)y & 7) == 0, DD_ERR_INVALID, "dd_recovery_fwd: y must be 8-byte aligned");
    const bool fast = !precise_mode();
    if (IcA) return fast ? launch_fwd2<true, true>(x, A, IcA, feat, y, B, H, W, st) : launch_fwd2<true, false>(x, A, IcA, feat, y, B, H, W, st);
    return fast ? launch_fwd2<false, true>(x, A, nullptr, feat, y, B, H, W, st) : launch_fwd2<false, false>(x, A, nullptr, feat, y, B, H, W, st);
}
