// dd_recovery_fwd.cu -- a6..a12: the fused filter chain forward (regressors, DeDark -> WB -> Gamma -> Contrast ->
// USM) in one pass over the image.
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,295-304,316-317.
//
// "Marching strips" (dd_layout.cuh): every (plane, 128-column strip) is cut into row-blocks of 32 rows; 296
// persistent CTAs (2 per SM, 256 threads) each own a contiguous, equal range of row-blocks and march down them:
//   1. stage   x0 -> pointwise chain -> x4 for 32 rows x (128+24) columns.  Global loads are 128-bit and issued one
//              block ahead (register prefetch); columns outside the image are not loaded at all -- the reflect
//              halo is filled by mirroring the freshly computed x4 values inside shared memory.  x4 is stored with
//              ROW PAIRS INTERLEAVED, XS2[row/2][col][row%2], so that                                -> ring XS2
//   2. H pass  pairs (row 2r, row 2r+1) of one column are aligned 64-bit operands: every FFMA2 of the 25-tap pass
//              takes its operand straight from an LDS.128 and the tap as an immediate (no shifted operand pairs).
//              Thread = 2 rows x 8 columns, 16 LDS.128, 200 FFMA2.                                    -> ring HS
//   3. V pass  row-major HS, pairs (col 2c, col 2c+1): thread = 2 columns x 8 rows, 32 LDS.64 with immediate offsets
//              (the first 24 ring rows are stored twice so a 32-row window never wraps), 200 FFMA2, then
//              y = (x4 - blur) p + x4 with 64-bit coalesced streaming stores.
// The vertical halo (24 rows) is paid once per CTA range.  The blur is FMA-pipe bound (50 FMA per pixel ~ 28 us per
// 16x3x640^2 at the measured 35 TFMA/s); FFMA2 halves its issue slots so loads, MUFU and index arithmetic issue in the
// shadow of the FMA pipe.
#include <cstring>

#include "dd_async.cuh"
#include "dd_recovery.cuh"
#include "dd_recovery_tc_fwd.cuh"

namespace dd {

constexpr int kPairs = 32;              // XS2 ring depth in row pairs (64 rows)
constexpr int kXP2 = 308;               // XS2 pitch (floats per row pair): 2*152 + 4, (kXP2/4) odd -> conflict-free LDS.128
constexpr int kHRows = kHRing + 24;     // HS rows: 64-row ring + the first 24 rows mirrored behind it
constexpr int kStage2 = (16 * kXW4 + kThreads - 1) / kThreads;  // (row pair, float4 column) items per thread per block (3)
constexpr int kXRP = 160;               // pitch (floats) of the TMA staging rows: 640 B, a multiple of the 128 B TMA alignment
// "mixed" schedule (the TMA instantiation, round 2): two barriers per row-block, the pointwise stage of block n+1 shares an interval
// with the horizontal pass of block n (half of the warps run them in the other order, so MUFU / issue-bound and FMA-bound code meet
// in time), the vertical pass of block n has the other interval -- the structure of the backward kernel.  It needs the x4 rows of
// block n+1 beside the centre rows the vertical pass of block n still reads: an 80-row x4 ring; the H ring gives its mirrored
// rows back (the vertical pass walks four 8-row groups that never wrap): 20 + 49.3 + 33.8 + 2.6 KB, still two CTAs per SM.
constexpr int kPairsM = 40;             // x4 ring of the mixed schedule, in row pairs

// TMA: the raw x rows of every row-block arrive by cp.async.bulk.tensor -- two 16-row x 160-column boxes for a block inside the
// image, one row box per (reflected) image row for the blocks at its top and bottom -- issued by warp 0 one block ahead into a
// staging tile and counted on an mbarrier: no register prefetch, no per-thread address arithmetic, and the stage phase never
// waits for DRAM.
// U8: x is the dataloader's uint8 batch (train.py:72) and `dark_tab` the 256-entry darkening table
// pow(k/255, p) (dd_dark_table): the darkened fp32 batch is never materialised in HBM, the stage phase looks it up.  Needs
// ALIGNED (W % 4 == 0: one 32-bit load per four pixels) and runs on the register-prefetch path (TMA = false).
template <bool HAS_ICA, bool FAST, bool ALIGNED, bool TMA, bool U8 = false>
__global__ void __launch_bounds__(kThreads, 2)
recovery_fwd_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap16,
                    const float* __restrict__ x, const float* __restrict__ A,
                    const float* __restrict__ IcA, const float* __restrict__ feat, float* __restrict__ y, int B, int H, int W,
                    const float* __restrict__ dark_tab = nullptr) {
    static_assert(!U8 || (ALIGNED && !TMA), "uint8 sources: aligned rows, register prefetch");
    pdl_launch();   // the wait follows the barrier set-up below: nothing in front of it touches global memory
    __shared__ float s_tab[U8 ? 256 : 1];
    const unsigned char* x8 = reinterpret_cast<const unsigned char*>(x);
    auto lut4 = [&](unsigned w) { return make_float4(s_tab[w & 255u], s_tab[(w >> 8) & 255u], s_tab[(w >> 16) & 255u], s_tab[w >> 24]); };
    extern __shared__ __align__(128) float smem[];
    constexpr bool MIXED = TMA;                        // see kPairsM
    constexpr int P = MIXED ? kPairsM : kPairs;        // x4 ring depth in row pairs
    constexpr int HR = MIXED ? kHRing : kHRows;        // H ring rows (mixed: no mirrored copy behind the ring)
    float* XR = smem;                                  // [32 rows][kXRP] raw x of the block being staged (TMA only)
    float* XS2 = smem + (TMA ? kRB * kXRP : 0);
    float* HS = XS2 + P * kXP2;
    float* MS = HS + HR * kHP;  // per virtual row: m = (1-c) + c*q
    auto xs2_row = [&](int pair) { return XS2 + (MIXED ? pair % P : (pair & (P - 1))) * kXP2; };
    __shared__ ImgParams sp;
    __shared__ __align__(8) uint64_t tma_bar;
    uint32_t tma_phase = 0;
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&tma_bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    // The grid dependency of this kernel is the predictor's output (feat).  The batch itself (x, IcA, A) is an input of the step:
    // the kernels of this library that write image-sized tensors (synthesis, dark-channel prior) never release their dependents
    // early (pdl_wait_only, dd_common.cuh), so whatever runs behind them sees them complete.  The first segment therefore requests
    // its per-row columns and its first tile BEFORE the wait, which sits in front of the regressors below.  (uint8 sources wait
    // here: the table.)
    bool waited = false;
    if (U8) {
        pdl_wait();
        waited = true;
        if (threadIdx.x < 256) s_tab[threadIdx.x] = __ldcg(dark_tab + threadIdx.x);   // not __ldg: ptxas hoists ld.global.nc above the wait
        __syncthreads();
    }

    const int tid = threadIdx.x;
    const Sched sc = make_sched(B, H, W);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const bool w2 = (W & 1) == 0;
    const int L = W - 1;

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        __syncthreads();  // previous segment fully consumed (rings, MS, sp)
        const float* xp = x + (size_t)u.plane * H * W;
        const unsigned char* xp8 = x8 + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
        float* yp = y + (size_t)u.plane * H * W;
        // columns 0..2 of every row of the segment (per-row contrast scalars): requested before the regressors are evaluated
        constexpr int kPer = (kMaxU + kThreads - 1) / kThreads;
        float x0r[kPer][3], icr[kPer][3];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int v = tid + k * kThreads;
            if (v < u.nU) {
                const int row = reflect(u.r0 - kRadius + v, H);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    x0r[k][c] = U8 ? s_tab[__ldg(xp8 + (size_t)row * W + c)] : __ldg(xp + (size_t)row * W + c);
                    icr[k][c] = HAS_ICA ? __ldg(ip + (size_t)row * W + c) : kDefaultIcA;
                }
            }
        }
        // item (k) of a thread: row pair rp (2 rows) x float4 column c4; fixed across blocks
        float4 pre[kStage2][2], prei[kStage2][2];
        auto stage = [&](int n) {
            if (TMA) {
                if (tid < 32) {
                    const int nrows = min(kRB, u.nU - n * kRB);
                    const int first = u.r0 - kRadius + n * kRB;  // image row of the block's first virtual row
                    if (tid == 0) mbar_arrive_expect_tx(&tma_bar, (uint32_t)nrows * kXRP * 4u);
                    __syncwarp();
                    if (nrows == kRB && first >= 0 && first + kRB <= H) {  // interior block: two 16-row boxes
                        if (tid < 2) tma_load_3d(XR + tid * 16 * kXRP, &xmap16, u.c0 - kRadius, first + 16 * tid, u.plane, &tma_bar);
                    } else if (tid < nrows) {                              // border block: one row box per reflected row
                        const int row = min(max(reflect(first + tid, H), 0), H - 1);
                        tma_load_3d(XR + tid * kXRP, &xmap, u.c0 - kRadius, row, u.plane, &tma_bar);
                    }
                }
                return;
            }
#pragma unroll
            for (int k = 0; k < kStage2; ++k) {
                const int f = tid + k * kThreads;
                const int rp = f / kXW4, c4 = f - rp * kXW4;
                const int v0 = n * kRB + 2 * rp;
                const int gc = u.c0 - kRadius + 4 * c4;
                if (f < 16 * kXW4 && v0 < u.nU) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = min(max(reflect(u.r0 - kRadius + v0 + e, H), 0), H - 1);
                        const size_t ro = (size_t)row * W;
                        if (ALIGNED) {
                            if (gc >= 0 && gc < W) {
                                if (U8) pre[k][e].x = __uint_as_float(__ldg(reinterpret_cast<const unsigned*>(xp8 + ro + gc)));  // four source bytes
                                else pre[k][e] = __ldg(reinterpret_cast<const float4*>(xp + ro + gc));
                                if (HAS_ICA) prei[k][e] = __ldg(reinterpret_cast<const float4*>(ip + ro + gc));
                            }
                        } else {
                            int g[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) g[i] = min(max(reflect(gc + i, W), 0), W - 1);
                            pre[k][e] = make_float4(__ldg(xp + ro + g[0]), __ldg(xp + ro + g[1]), __ldg(xp + ro + g[2]), __ldg(xp + ro + g[3]));
                            if (HAS_ICA)
                                prei[k][e] = make_float4(__ldg(ip + ro + g[0]), __ldg(ip + ro + g[1]), __ldg(ip + ro + g[2]), __ldg(ip + ro + g[3]));
                        }
                    }
                }
            }
        };
        if (TMA) stage(0);  // the first block's rows fly while the regressors and the per-row scalars below are computed
        if (!waited) {
            pdl_wait();
            waited = true;
        }
        if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
        __syncthreads();
        const ChainK ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
        const float pc = sp.c, pp = sp.p;
        {   // per-row contrast scalars of the segment (the rows' first three columns were loaded at the top of the segment)
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int v = tid + k * kThreads;
                if (v < u.nU) {
                    float x3[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x3[c] = chain_x3<HAS_ICA, FAST>(ck, x0r[k][c], icr[k][c]);
                    const RowLum rl = row_lum<true>(x3[0], x3[1], x3[2]);
                    MS[v] = (1.f - pc) + pc * rl.q;
                }
            }
        }

        if (!TMA) stage(0);

        auto stage_compute = [&](int n) {   // raw x rows of block n (staging tile / prefetch registers) -> pointwise chain -> x4 ring
#pragma unroll
            for (int k = 0; k < kStage2; ++k) {
                const int f = tid + k * kThreads;
                const int rp = f / kXW4, c4 = f - rp * kXW4;
                const int v0 = n * kRB + 2 * rp;
                const int gc = u.c0 - kRadius + 4 * c4;
                if (f < 16 * kXW4 && v0 < u.nU && (!ALIGNED || (gc >= 0 && gc < W))) {
                    float o[2][4];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float m = MS[min(v0 + e, u.nU - 1)];
                        const float4 in = TMA ? *reinterpret_cast<const float4*>(XR + (2 * rp + e) * kXRP + 4 * c4)
                                          : U8 ? lut4(__float_as_uint(pre[k][e].x)) : pre[k][e];
                        const float4 ic = HAS_ICA ? prei[k][e] : make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                        o[e][0] = chain_x3<HAS_ICA, FAST>(ck, in.x, ic.x) * m;
                        o[e][1] = chain_x3<HAS_ICA, FAST>(ck, in.y, ic.y) * m;
                        o[e][2] = chain_x3<HAS_ICA, FAST>(ck, in.z, ic.z) * m;
                        o[e][3] = chain_x3<HAS_ICA, FAST>(ck, in.w, ic.w) * m;
                    }
                    float* xrow = xs2_row(v0 >> 1);  // element (row e, staged col c) at xrow[2c + e]
                    float4* dst = reinterpret_cast<float4*>(xrow + 8 * c4);
                    dst[0] = make_float4(o[0][0], o[1][0], o[0][1], o[1][1]);
                    dst[1] = make_float4(o[0][2], o[1][2], o[0][3], o[1][3]);
                    if (ALIGNED) {
                        // reflect halo (F.pad mode='reflect', filtersB.py:167): image col -j <- j, L+d <- L-d
                        if (gc <= kRadius) {  // only strip 0 has image columns 0..12 at gc <= 12
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int j = gc + i;
                                if (j >= 1 && j <= kRadius)
                                    *reinterpret_cast<float2*>(xrow + 2 * (kRadius - j - u.c0)) = make_float2(o[0][i], o[1][i]);
                            }
                        }
                        if (gc + 3 >= L - kRadius) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int d = L - (gc + i);
                                const int tc = L + d - u.c0 + kRadius;
                                if (d >= 1 && d <= kRadius && tc < kXW)
                                    *reinterpret_cast<float2*>(xrow + 2 * tc) = make_float2(o[0][i], o[1][i]);
                            }
                        }
                    }
                }
            }
        };
        auto hpass = [&](int n) {
            {   // horizontal pass: thread = row pair x 8 columns; lanes 0..15 -> row pairs
                const int rp = tid & 15, cg = tid >> 4;
                const int v0 = n * kRB + 2 * rp;
                u64 acc[8];
                blur8_pairs<2>(xs2_row(v0 >> 1) + 16 * cg, acc);
                float lo[8], hi[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float2 a = upk(acc[t]);
                    lo[t] = a.x;
                    hi[t] = a.y;
                }
                const int s0 = v0 & (kHRing - 1);  // even; rows s0, s0 + 1
                float* h0 = HS + s0 * kHP + 8 * cg;
                *reinterpret_cast<float4*>(h0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(h0 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                *reinterpret_cast<float4*>(h0 + kHP) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(h0 + kHP + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                if (!MIXED && s0 < 24) {  // mirrored copy behind the ring: a 32-row window starting at slot <= 56 never wraps
                    float* h1 = h0 + kHRing * kHP;
                    *reinterpret_cast<float4*>(h1) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<float4*>(h1 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                    *reinterpret_cast<float4*>(h1 + kHP) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(h1 + kHP + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                }
            }
        };
        auto vpass = [&](int n) {
            {   // vertical pass + USM epilogue: lanes -> column pairs, 8 rows per thread
                const int col2 = 2 * (tid & 63), rg = tid >> 6;
                const int o_first = n * kRB - kRadius + 8 * rg;  // virtual row of the first output (even)
                if (o_first >= kRadius && o_first < kRadius + u.seg_len) {
                    u64 bl[8];
                    if (MIXED) vpass8x2(HS, (o_first - kRadius) & (kHRing - 1), col2, bl);   // four 8-row groups, no mirrored rows
                    else blur8_pairs<kHP>(HS + ((o_first - kRadius) & (kHRing - 1)) * kHP + col2, bl);
                    const int gc = u.c0 + col2;
                    if (gc < W) {
                        const u64 p2 = pk(pp, pp), m1 = pk(-1.f, -1.f);
                        float* drow = yp + (size_t)(u.r0 + o_first - kRadius) * W + gc;
                        const int rows_left = kRadius + u.seg_len - o_first;  // >= 1
#pragma unroll
                        for (int r = 0; r < 8; r += 2) {
                            // centre x4 of rows (o, o+1), columns (c, c+1): one LDS.128 from the interleaved ring
                            const float4 c4v = *reinterpret_cast<const float4*>(xs2_row((o_first + r) >> 1) + 2 * (col2 + kRadius));
                            const u64 xa = pk(c4v.x, c4v.z), xb = pk(c4v.y, c4v.w);
                            const u64 ya = fma2(fma2(bl[r], m1, xa), p2, xa);      // (x4 - blur) * p + x4
                            const u64 yb = fma2(fma2(bl[r + 1], m1, xb), p2, xb);
                            if (w2) {
                                if (r < rows_left) __stcs(reinterpret_cast<float2*>(drow + (size_t)r * W), upk(ya));
                                if (r + 1 < rows_left) __stcs(reinterpret_cast<float2*>(drow + (size_t)(r + 1) * W), upk(yb));
                            } else {
                                const float2 ta = upk(ya), tb = upk(yb);
                                if (r < rows_left) {
                                    drow[(size_t)r * W] = ta.x;
                                    if (gc + 1 < W) drow[(size_t)r * W + 1] = ta.y;
                                }
                                if (r + 1 < rows_left) {
                                    drow[(size_t)(r + 1) * W] = tb.x;
                                    if (gc + 1 < W) drow[(size_t)(r + 1) * W + 1] = tb.y;
                                }
                            }
                        }
                    }
                }
            }
        };

        if (MIXED) {
            // two barriers per row-block: [H pass of block n | stage of block n+1] , [V pass of block n]
            __syncthreads();                       // MS ready; the rings of the previous segment are no longer read
            mbar_wait(&tma_bar, tma_phase);        // block 0's rows (requested before the per-row scalars)
            tma_phase ^= 1u;
            stage_compute(0);
            __syncthreads();                       // x4 rows of block 0 visible, staging tile consumed
            if (u.nB > 1) stage(1);
            const bool stage_first = (tid >> 7) & 1;   // warps 4..7 stage first, warps 0..3 blur first: one of each on every scheduler (warp % 4)
            for (int n = 0; n < u.nB; ++n) {
                const bool more = n + 1 < u.nB;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    if ((pass == 0) == stage_first) {
                        if (more) {
                            mbar_wait(&tma_bar, tma_phase);
                            tma_phase ^= 1u;
                            stage_compute(n + 1);
                        }
                    } else {
                        hpass(n);
                    }
                }
                __syncthreads();                   // H rows of block n and x4 rows of block n+1 visible; staging tile consumed
                if (n + 2 < u.nB) stage(n + 2);
                vpass(n);
                __syncthreads();                   // the centre rows block n+2's stage overwrites and the H rows block n+1 overwrites are free
            }
        } else {
            for (int n = 0; n < u.nB; ++n) {
                __syncthreads();  // MS ready (n == 0); ring slots of block n no longer read by the previous V pass
                stage_compute(n);
                if (n + 1 < u.nB) stage(n + 1);
                __syncthreads();
                hpass(n);
                __syncthreads();
                vpass(n);
            }
        }
    }
}

constexpr size_t kFwdSmem = (size_t)(kPairs * kXP2 + kHRows * kHP + kMaxU) * sizeof(float);
constexpr size_t kFwdSmemTma = (size_t)(kRB * kXRP + kPairsM * kXP2 + kHRing * kHP + kMaxU) * sizeof(float);   // mixed schedule (kPairsM)

template <bool HAS_ICA, bool FAST, bool ALIGNED, bool TMA>
static int launch_fwd4(const CUtensorMap& xmap, const CUtensorMap& xmap16, const float* x, const float* A, const float* IcA, const float* feat, float* y, int B,
                       int H, int W, const Sched& sc, cudaStream_t st) {
    constexpr size_t smem = TMA ? kFwdSmemTma : kFwdSmem;
    DD_ENSURE_SMEM((recovery_fwd_kernel<HAS_ICA, FAST, ALIGNED, TMA>), smem, "recovery kernel");
    launch_pdl(recovery_fwd_kernel<HAS_ICA, FAST, ALIGNED, TMA>, dim3(sc.G), dim3(kThreads), smem, st, xmap, xmap16, x, A, IcA, feat, y, B, H, W,
               (const float*)nullptr);
    return DD_OK;
}

// the same kernel reading the uint8 batch through the darkening table
template <bool HAS_ICA, bool FAST>
static int launch_fwd_u8(const uint8_t* src, const float* tab, const float* A, const float* IcA, const float* feat, float* y, int B, int H, int W,
                         cudaStream_t st) {
    const Sched sc = make_sched(B, H, W);
    CUtensorMap none;
    memset(&none, 0, sizeof(none));
    auto kern = recovery_fwd_kernel<HAS_ICA, FAST, true, false, true>;
    DD_ENSURE_SMEM(kern, kFwdSmem, "recovery kernel (uint8 source)");
    launch_pdl(kern, dim3(sc.G), dim3(kThreads), kFwdSmem, st, none, none, reinterpret_cast<const float*>(src), A, IcA, feat, y, B, H, W, tab);
    count_launch();
    return check_launch("dd_recovery_fwd_u8");
}

// tensor-core variant: 148 persistent CTAs, 3xTF32 (fp32 gate 1e-5)
template <bool HAS_ICA, bool FAST, bool DBG>
static int launch_fwd_tc(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H, int W,
                         bool x3, cudaStream_t st) {
    constexpr int R = 48;
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    if (x3) {
        auto kern = btc::recovery_fwd_tc_kernel<R, true, HAS_ICA, FAST, DBG>;
        constexpr size_t smem = btc::Lay<R, true>::SMEM;
        DD_ENSURE_SMEM(kern, smem, "recovery_fwd_tc_kernel");
        launch_pdl(kern, dim3(sc.G), dim3(btc::kThreadsTC), smem, st, x, A, IcA, feat, y, B, H, W);
    } else {
        auto kern = btc::recovery_fwd_tc_kernel<R, false, HAS_ICA, FAST, DBG>;
        constexpr size_t smem = btc::Lay<R, false>::SMEM;
        DD_ENSURE_SMEM(kern, smem, "recovery_fwd_tc_kernel (1xTF32)");
        launch_pdl(kern, dim3(sc.G), dim3(btc::kThreadsTC), smem, st, x, A, IcA, feat, y, B, H, W);
    }
    count_launch();
    return check_launch("dd_recovery_fwd (tensor-core blur)");
}

template <bool HAS_ICA, bool FAST, bool ALIGNED>
static int launch_fwd3(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H,
                       int W, cudaStream_t st) {
    if (ALIGNED && blur_on_tensor_cores()) return launch_fwd_tc<HAS_ICA, FAST, false>(x, A, IcA, feat, y, B, H, W, true, st);
    const Sched sc = make_sched(B, H, W);
    CUtensorMap xmap, xmap16;
    memset(&xmap, 0, sizeof(xmap));
    memset(&xmap16, 0, sizeof(xmap16));
    // TMA staging: 16-byte aligned rows (ALIGNED), the default IcA (no second tile to stage) and an image at least one
    // row box wide; otherwise the register-prefetch path
    const bool tma = ALIGNED && !HAS_ICA && W >= kXRP && H >= 16 && make_tensor_map_3d(&xmap, x, B * 3, H, W, kXRP, 1) &&
                     make_tensor_map_3d(&xmap16, x, B * 3, H, W, kXRP, 16);
    if (int e = tma ? launch_fwd4<HAS_ICA, FAST, ALIGNED, ALIGNED && !HAS_ICA>(xmap, xmap16, x, A, IcA, feat, y, B, H, W, sc, st)
                    : launch_fwd4<HAS_ICA, FAST, ALIGNED, false>(xmap, xmap16, x, A, IcA, feat, y, B, H, W, sc, st))
        return e;
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

namespace dd {
// bf16 I/O mode (the reference trains under autocast, engine/trainer.py:330): x and / or y are bf16, the
// blur runs as plain TF32 on the tensor cores (no operand split: bf16 values are exact in TF32), fp32 arithmetic in between.
template <bool HAS_ICA, typename TX, typename TY>
static int launch_fwd_tc_io(const TX* x, const float* A, const float* IcA, const float* feat, TY* y, int B, int H, int W, cudaStream_t st) {
    constexpr int R = 48;
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    auto kern = btc::recovery_fwd_tc_kernel<R, false, HAS_ICA, true, false, TX, TY>;
    constexpr size_t smem = btc::Lay<R, false>::SMEM;
    DD_ENSURE_SMEM(kern, smem, "recovery_fwd_tc_kernel (bf16 I/O)");
    launch_pdl(kern, dim3(sc.G), dim3(btc::kThreadsTC), smem, st, x, A, IcA, feat, y, B, H, W);
    count_launch();
    return check_launch("dd_recovery_fwd_ex");
}
}  // namespace dd

extern "C" int dd_recovery_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y, int B, int H, int W, void* stream_);

extern "C" int dd_recovery_fwd_ex(const void* x, int x_dtype, const float* A, const float* IcA, const float* feat, void* y, int y_dtype,
                                  int B, int H, int W, void* stream_) {
    using namespace dd;
    typedef __nv_bfloat16 bf16;
    if (x_dtype == DD_F32 && y_dtype == DD_F32)
        return dd_recovery_fwd(reinterpret_cast<const float*>(x), A, IcA, feat, reinterpret_cast<float*>(y), B, H, W, stream_);
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_fwd_ex", B, H, W)) return e;
    DD_REQUIRE((x_dtype == DD_F32 || x_dtype == DD_BF16) && (y_dtype == DD_F32 || y_dtype == DD_BF16), DD_ERR_INVALID, "dd_recovery_fwd_ex: unknown dtype");
    DD_REQUIRE(x && feat && y && x != y, DD_ERR_INVALID, "dd_recovery_fwd_ex: null pointer or y aliases x");
    DD_REQUIRE((W & 3) == 0 && ((uintptr_t)x & 15) == 0 && (!IcA || ((uintptr_t)IcA & 15) == 0), DD_ERR_INVALID,
               "dd_recovery_fwd_ex: the bf16 I/O mode needs W %% 4 == 0 and 16-byte aligned x / IcA (got W = %d)", W);
#define DD_FWD_IO(TX, TY)                                                                                                          \
    return IcA ? launch_fwd_tc_io<true, TX, TY>(reinterpret_cast<const TX*>(x), A, IcA, feat, reinterpret_cast<TY*>(y), B, H, W, st) \
               : launch_fwd_tc_io<false, TX, TY>(reinterpret_cast<const TX*>(x), A, nullptr, feat, reinterpret_cast<TY*>(y), B, H, W, st)
    if (x_dtype == DD_BF16 && y_dtype == DD_BF16) { DD_FWD_IO(bf16, bf16); }
    if (x_dtype == DD_BF16) { DD_FWD_IO(bf16, float); }
    DD_FWD_IO(float, bf16);
#undef DD_FWD_IO
}

// unit-test hook: the bare reflect-padded 25x25 Gaussian (filtersB.py:154-175: F.pad reflect + conv2d) of x through the
// tensor-core engine, x3 != 0: 3xTF32, else 1xTF32.  Needs W % 4 == 0 and 16-byte aligned x.
extern "C" int dd_debug_blur_tc(const float* x, float* y, int B, int H, int W, int x3, void* stream_) {
    using namespace dd;
    if (int e = check_recovery_shape("dd_debug_blur_tc", B, H, W)) return e;
    DD_REQUIRE(x && y && x != y && (W & 3) == 0 && ((uintptr_t)x & 15) == 0, DD_ERR_INVALID, "dd_debug_blur_tc: W %% 4 == 0 and aligned x required");
    return launch_fwd_tc<false, true, true>(x, nullptr, nullptr, x /* feat: unused */, y, B, H, W, x3 != 0, (cudaStream_t)stream_);
}

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

// ---- uint8 batch -> filter chain without a darkened fp32 batch in HBM -------------------------------------
extern "C" int dd_recovery_fwd_u8(const uint8_t* src, const float* dark_table, const float* A, const float* IcA, const float* feat, float* y,
                                  int B, int H, int W, void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    if (int e = check_recovery_shape("dd_recovery_fwd_u8", B, H, W)) return e;
    DD_REQUIRE(src && dark_table && feat && y, DD_ERR_INVALID, "dd_recovery_fwd_u8: null pointer");
    DD_REQUIRE((W & 3) == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)y & 7) == 0 && (!IcA || ((uintptr_t)IcA & 15) == 0), DD_ERR_INVALID,
               "dd_recovery_fwd_u8: needs W %% 4 == 0, a 4-byte aligned source and 16-byte aligned IcA (got W = %d)", W);
    const bool fast = !precise_mode();
    if (IcA) return fast ? launch_fwd_u8<true, true>(src, dark_table, A, IcA, feat, y, B, H, W, st) : launch_fwd_u8<true, false>(src, dark_table, A, IcA, feat, y, B, H, W, st);
    return fast ? launch_fwd_u8<false, true>(src, dark_table, A, nullptr, feat, y, B, H, W, st) : launch_fwd_u8<false, false>(src, dark_table, A, nullptr, feat, y, B, H, W, st);
}
