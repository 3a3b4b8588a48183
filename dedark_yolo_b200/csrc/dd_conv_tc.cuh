// dd_conv_tc.cuh -- the predictor's dense conv GEMMs (layers 2..5 of ExtractParameters2, nn/modules/common.py:52-78)
// on the 5th-generation tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM, operands staged in shared
// memory in the canonical no-swizzle core-matrix layouts, one elected thread issuing the MMAs, completion through
// tcgen05.commit -> mbarrier, epilogue through tcgen05.ld.
//
// Precision.  The 1e-5 output gate of the module does not survive TF32 operand rounding, so every operand is split
// on the fly into  x = hi + lo  (hi = rna_tf32(x), lo = rna_tf32(x - hi); both exactly representable in TF32) and
// every product is issued as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM ("3xTF32"; the first two share one
// MMA whose B operand is the stacked [hi | lo], the epilogue adds the two accumulator column blocks): the dropped
// lo*lo term and the rounding of lo are each <= 2^-22 relative.
//
// Three GEMMs per layer (3x3 / stride 2 / pad 1 conv, Cin -> Cout, input HIN x HIN, output HO = HIN/2):
//   forward  D[pixel][co]        = sum_{tap,ci}  im2col[pixel][tap,ci] * W[co][ci][tap]            (M=128, N=Cout)
//   dgrad    D[quad][ci,parity]  = sum_{co,nb}   dOut[co][quad + nb]   * Wq[ci,parity][co,nb]      (M=128, N=4*Cin)
//            a "quad" (a,c) is the 2x2 block of input pixels (2a+py, 2c+px); nb = (da,dc) in {0,1}^2 runs over
//            the 2x2 neighbourhood of OUTPUT pixels (a+da, c+dc) it touches; Wq holds w[kh][kw] with
//            kh = py - 2*da + 1, kw = px - 2*dc + 1 where that is a valid tap and 0 elsewhere
//   wgrad    D[tap,ci | 1][co]   = sum_{pixel}   im2col^T[tap,ci | 1][pixel] * dOut^T[co][pixel]   (M=128 blocks, N=Cout)
//            (both operands K-major with K = pixels: 16-byte chunks hold 4 consecutive pixels of one row; the extra
//             row of ones yields the bias gradient)
// M index = 128 consecutive pixels (or quads) in (b, row, col) raster order, so any batch size works.
#pragma once
#include "dd_async.cuh"
#include "dd_common.cuh"
#include "dd_layout.cuh"
#include "dd_tcgen05.cuh"

namespace dd {
namespace tc {

// optional in-kernel time stamps (profiles/microbench/tc_probe.cu defines DD_TC_TIMING); no code otherwise
#ifdef DD_TC_TIMING
__device__ long long g_tc_stamp[64];
#define DD_TC_STAMP(i)                                                         \
    do {                                                                       \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_stamp[(i)] = clock64();  \
    } while (0)
__device__ long long g_tc_cta_cycles[256];
#define DD_TC_CTA_BEGIN() const long long cta_t0__ = clock64()
#define DD_TC_CTA_END()                                                        \
    do {                                                                       \
        if (threadIdx.x == 0) g_tc_cta_cycles[blockIdx.x] = clock64() - cta_t0__; \
    } while (0)
#define DD_TC_STAMP_C(i, thr, ctaidx)                                              \
    do {                                                                          \
        if ((ctaidx) == 0 && threadIdx.x == (thr)) g_tc_stamp[(i)] = clock64();    \
    } while (0)
#define DD_TC_STAMP_T(i, thr)                                                     \
    do {                                                                         \
        if (blockIdx.x == 0 && threadIdx.x == (thr)) g_tc_stamp[(i)] = clock64(); \
    } while (0)
#else
#define DD_TC_STAMP(i) \
    do {               \
    } while (0)
#define DD_TC_STAMP_T(i, thr) \
    do {                      \
    } while (0)
#define DD_TC_STAMP_C(i, thr, ctaidx) \
    do {                              \
    } while (0)
#define DD_TC_CTA_BEGIN() \
    do {                  \
    } while (0)
#define DD_TC_CTA_END() \
    do {                \
    } while (0)
#endif

// ---- prepared weights (one launch per forward; the weights change every optimizer step) ----------------------
// per layer l = 1..4 (conv2..conv5), floats:
//   fwd  : [stage c = ci/8][slot s = 9*((ci%8)/4) + tap][row = co (hi) | COUT + co (lo)][e = ci%4] = 2 * 9*CIN*COUT
//   dgrad: [stage c = co/8][slot s = co%8][row = n (hi) | 4*CIN + n (lo); n = ci*4 + py*2 + px][e = da*2 + dc] = 2 * 16*CIN*COUT
__host__ __device__ constexpr int prep_fwd_elems(int cin, int cout) { return 2 * 9 * cin * cout; }
__host__ __device__ constexpr int prep_dgrad_elems(int cin, int cout) { return 2 * 16 * cin * cout; }

struct PrepJob {
    const float* w;  // [COUT][CIN][3][3]
    float* fwd;
    float* dgrad;
    int cin, cout;
};
struct PrepJobs {
    PrepJob j[4];
};

constexpr int kPrepBlocksPerJob = 32;  // CTAs (of 256 threads) per layer; 4 layers
__device__ __forceinline__ void prep_weights_body(const PrepJobs& jobs, const int vblock) {
    const PrepJob jb = jobs.j[vblock / kPrepBlocksPerJob];
    const int cin = jb.cin, cout = jb.cout;
    const int nf = 9 * cin * cout, nd = 16 * cin * cout;  // hi elements of each table
    for (int i = (vblock % kPrepBlocksPerJob) * 256 + threadIdx.x; i < nf + nd; i += kPrepBlocksPerJob * 256) {
        float v;
        float *hi, *lo;
        if (i < nf) {
            const int e = i & 3, co = (i >> 2) % cout, s = (i / (4 * cout)) % 18, c = i / (4 * cout * 18);
            const int tap = s % 9, ci = 8 * c + 4 * (s / 9) + e;
            v = __ldg(jb.w + ((size_t)co * cin + ci) * 9 + tap);
            hi = jb.fwd + (size_t)c * (18 * 2 * cout * 4) + (size_t)(s * 2 * cout + co) * 4 + e;
            lo = hi + cout * 4;
        } else {
            const int k = i - nf, N = 4 * cin;
            const int e = k & 3, n = (k >> 2) % N, s = (k / (4 * N)) % 8, c = k / (4 * N * 8);
            const int co = 8 * c + s, ci = n >> 2, py = (n >> 1) & 1, px = n & 1, da = e >> 1, dc = e & 1;
            const int kh = py - 2 * da + 1, kw = px - 2 * dc + 1;
            v = (kh >= 0 && kw >= 0) ? __ldg(jb.w + ((size_t)co * cin + ci) * 9 + kh * 3 + kw) : 0.f;  // kh, kw <= 2 always
            hi = jb.dgrad + (size_t)c * (8 * 2 * N * 4) + (size_t)(s * 2 * N + n) * 4 + e;
            lo = hi + N * 4;
        }
        const float h = tf32_rna(v);
        *hi = h;
        *lo = tf32_rna(v - h);
    }
}

// =============================================================================================================
// forward:  out = leaky(conv(in, W) + bias)
//   Persistent, warp-specialised CTA (one per SM): warps 0..7 produce, warp 8 issues the MMAs.
//   * the prepared weights (B operand) of ALL K stages are brought into shared memory once per CTA (cp.async);
//   * the im2col A operand never touches shared memory: a tile = 128 output pixels = the 128 TMEM lanes; producer
//     thread (pixel m = t%128, half h = t/128) gathers the 3x3 windows of 4 channels (one stage = 8 channels = 72 k),
//     splits them and writes its 36 hi + 36 lo values straight into its TMEM lane (tcgen05.st), into one of two A
//     buffers; the MMAs take A from tensor memory (with N = 32..64 an A operand in shared memory costs more
//     shared-memory bandwidth than the MMA costs tensor-pipe time: measured 94 cycles per MMA).  full[]/empty[]
//     mbarriers hand the buffers to the MMA warp and back (tcgen05.commit);
//   * per k-step two MMAs:  a_hi x [b_hi | b_lo] (N = 64)  and  a_lo x b_hi (N = 32, accumulating onto columns 0..31);
//     the epilogue adds the two column blocks, bias, LeakyReLU.
//   TMEM columns: [0,64) accumulator, then two A buffers of [72 hi | 72 lo].
// =============================================================================================================
template <int CIN, int COUT>
constexpr size_t conv_tc_fwd_smem() { return (size_t)((CIN / 8) * 18 * 2 * COUT * 4) * sizeof(float); }

struct FwdCtl {
    uint64_t full[2], empty[2], done;
    uint32_t tmem_base;
};

template <int CIN, int COUT, int HIN>
__global__ void __launch_bounds__(288, 1)
conv_tc_fwd(const float* __restrict__ in, const float* __restrict__ wprep, const float* __restrict__ bias,
            float* __restrict__ out, int total_px) {
    constexpr int HO = HIN / 2, NST = CIN / 8;
    constexpr int B_STAGE = 18 * 2 * COUT * 4;      // floats: [18 slots][hi rows | lo rows][4]
    constexpr uint32_t TMEM_COLS = 512, A_COL0 = 64, A_BUF = 144, A_LO = 72;
    static_assert(COUT == 32, "epilogue assumes 32 output channels");
    extern __shared__ __align__(128) float smem_tc[];
    __shared__ FwdCtl ctl;
    float* sB = smem_tc;  // [NST][18 slots][2*COUT rows][4]
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int ntiles = (total_px + 127) / 128;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    DD_TC_STAMP(0);
    if (warp == 8) tmem_alloc(&ctl.tmem_base, TMEM_COLS);
    if (t == 0) {
        mbar_init(&ctl.full[0], 256); mbar_init(&ctl.full[1], 256);
        mbar_init(&ctl.empty[0], 1); mbar_init(&ctl.empty[1], 1);
        mbar_init(&ctl.done, 1);
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // The prepared weights are written by conv1_fwd_prep_kernel at the head of the forward.  For conv3..conv5 (CIN == 32) that
    // kernel is at least the predecessor's predecessor, and every tensor-core kernel releases its dependents only AFTER its own
    // wait has returned (below: wait, then launch), so when this CTA runs, the predecessor is past its wait and the weights are
    // final: their 74 KB copy starts ahead of the grid dependency.  conv2 (CIN == 16) directly follows the producer and waits.
    constexpr bool EARLY_W = CIN == 32;
    if (EARLY_W && warp < 8) {
        for (int i = t; i < NST * B_STAGE / 4; i += 256) cp_async16(sB + 4 * i, wprep + 4 * i);
        cp_async_commit();
    }
    pdl_wait();
    pdl_launch();   // after the TMEM allocation (dd_common.cuh) and after the wait (see above)
    const uint32_t tmem = ctl.tmem_base;
    DD_TC_STAMP(1);

    if (warp < 8) {
        // ================================ producers + epilogue ================================
        const int m = t & 127, h = t >> 7, q = warp & 3;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        if (!EARLY_W) {
            for (int i = t; i < NST * B_STAGE / 4; i += 256) cp_async16(sB + 4 * i, wprep + 4 * i);
            cp_async_commit();
        }

        // im2col gather of one stage (4 channels x 3x3 window) into registers: per window row one aligned 64-bit load
        // (columns 2*ow, 2*ow+1); the left neighbour (column 2*ow-1) comes from the previous lane's load
        struct Raw {
            float2 x[12];  // [e*3 + kh]: columns 2*ow, 2*ow+1
            float own[12]; // column 2*ow-1, loaded by lane 0 only
        };
        auto gather = [&](Raw& v, int tile, int c) {
            const int gp = tile * 128 + m;
            const bool valid = gp < total_px;
            const int b = gp / (HO * HO), rem = gp % (HO * HO), oh = rem / HO, ow = rem % HO;
            const float* img = in + ((size_t)b * CIN + 8 * c + 4 * h) * HIN * HIN + 2 * ow;
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int ih = 2 * oh - 1 + kh;
                    const bool ok = valid && ih >= 0;
                    const float* rp = img + (size_t)e * HIN * HIN + ih * HIN;
                    v.x[e * 3 + kh] = ok ? __ldg(reinterpret_cast<const float2*>(rp)) : make_float2(0.f, 0.f);
                    v.own[e * 3 + kh] = (ok && lane == 0 && ow > 0) ? __ldg(rp - 1) : 0.f;
                }
        };
        // stage (it, c) + 2 in this CTA's sequence of stages, if there is one
        auto gather_ahead = [&](Raw& v, int it, int c) {
            c += 2;
            if (c >= NST) { c -= NST; ++it; }
            if (it < my_tiles) gather(v, blockIdx.x + it * gridDim.x, c);
        };
        int g = 0;  // running stage counter (A buffer = g & 1); NST is even, so g & 1 == c & 1
        auto stage = [&](Raw& v, int it, int c) {
            const int s = g & 1;
            if (g >= 2) mbar_wait(&ctl.empty[s], (uint32_t)((g >> 1) - 1) & 1u);  // stage g-2's MMAs have read buffer s
            fence_after_sync();
            DD_TC_STAMP(2 + 4 * (g & 7));
            const bool first_col = (m % HO) == 0;  // ow == 0: HO divides 128, so the column does not depend on the tile
            // k order within the stage: (half h, tap j, channel e) -> TMEM column h*36 + j*4 + e
            float hi[36], lo[36];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float nb = __shfl_up_sync(0xffffffffu, v.x[e * 3 + kh].y, 1);
                    const float left = first_col ? 0.f : (lane == 0 ? v.own[e * 3 + kh] : nb);
                    const float val[3] = {left, v.x[e * 3 + kh].x, v.x[e * 3 + kh].y};
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int k = (kh * 3 + kw) * 4 + e;
                        hi[k] = tf32_rna(val[kw]);
                        lo[k] = tf32_rna(val[kw] - hi[k]);
                    }
                }
            const uint32_t a_col = tmem + lane_base + A_COL0 + (uint32_t)s * A_BUF + (uint32_t)h * 36u;
            tmem_st16(a_col, hi); tmem_st16(a_col + 16, hi + 16); tmem_st4(a_col + 32, hi + 32);
            tmem_st16(a_col + A_LO, lo); tmem_st16(a_col + A_LO + 16, lo + 16); tmem_st4(a_col + A_LO + 32, lo + 32);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(&ctl.full[s]);
            DD_TC_STAMP(3 + 4 * (g & 7));
            gather_ahead(v, it, c);
            DD_TC_STAMP(4 + 4 * (g & 7));
            ++g;
        };
        static_assert(NST % 2 == 0, "two register stages alternate");
        Raw v0, v1;
        if (my_tiles > 0) {
            gather(v0, blockIdx.x, 0);
            gather(v1, blockIdx.x, 1);
        }
        cp_async_wait_all();
        fence_proxy_async();  // the weights in shared memory are read by the tensor core (async proxy)
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
#pragma unroll 1
            for (int c = 0; c < NST; c += 2) {
                stage(v0, it, c);
                stage(v1, it, c + 1);
            }
            // epilogue of this tile (the next tile's gathers are already in flight)
            mbar_wait(&ctl.done, (uint32_t)it & 1u);
            fence_after_sync();
            DD_TC_STAMP(40);
            const int half = warp >> 2;
            float r0[16], r1[16];
            tmem_ld16(tmem + lane_base + (uint32_t)(half * 16), r0);
            tmem_ld16(tmem + lane_base + (uint32_t)(COUT + half * 16), r1);
            fence_before_sync();  // orders the TMEM reads before the next full[] arrival (the next tile overwrites them)
            const int gpe = tile * 128 + q * 32 + lane;
            if (gpe < total_px) {
                const int be = gpe / (HO * HO), reme = gpe % (HO * HO);
                float* o = out + ((size_t)be * COUT + half * 16) * (HO * HO) + reme;
#pragma unroll
                for (int i = 0; i < 16; ++i) o[(size_t)i * (HO * HO)] = leaky((r0[i] + r1[i]) + __ldg(bias + half * 16 + i));
            }
            DD_TC_STAMP(41);
        }
    } else {
        // ================================ MMA issuer (warp 8) ================================
        constexpr uint32_t idesc64 = make_idesc(128, 2 * COUT, 0, 0), idesc32 = make_idesc(128, COUT, 0, 0);
        const uint32_t b_base = smem_u32(sB);
        const int nstages = my_tiles * NST;
        for (int g = 0; g < nstages; ++g) {
            const int s = g & 1, c = g % NST;
            mbar_wait(&ctl.full[s], (uint32_t)(g >> 1) & 1u);
            fence_after_sync();
            DD_TC_STAMP_T(44 + 2 * (g & 7), 256);
            if (elect_one()) {
                const uint32_t a0 = tmem + A_COL0 + (uint32_t)s * A_BUF, b0 = b_base + (uint32_t)c * (B_STAGE * 4u);
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const uint64_t bd = make_desc(b0 + (uint32_t)(2 * j) * (2 * COUT * 16u), 2 * COUT * 16u, 128u);
                    mma_tf32_ts(tmem, a0 + 8u * j, bd, idesc64, (c > 0 || j > 0) ? 1u : 0u);   // hi x [hi | lo]
                    mma_tf32_ts(tmem, a0 + A_LO + 8u * j, bd, idesc32, 1u);                     // lo x hi
                }
                mma_commit(&ctl.empty[s]);
                if (c == NST - 1) mma_commit(&ctl.done);
            }
            DD_TC_STAMP_T(45 + 2 * (g & 7), 256);
            __syncwarp();
        }
    }
    DD_TC_STAMP(42);
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, TMEM_COLS);
}

// =============================================================================================================
// backward of one layer in ONE launch of persistent, warp-specialised CTAs (one per SM, 288 threads: warps 0..7
// produce operands and run the epilogue, warp 8 issues the MMAs): the first n_w CTAs accumulate weight-gradient
// slices, the others compute the data gradient (the two are independent given dpre).
// =============================================================================================================
struct BwdCtl {
    uint64_t full[2], empty[2], done[2], acc_empty[2];
    uint32_t tmem_base;
};
constexpr uint32_t kBwdTmemCols = 512;

// -------------------------------------------------------------------------------------------------------------
// data gradient:  din = leaky'(act_in) * conv_transpose(dpre, W)      (gradient w.r.t. the previous PRE-activation)
//   tile = 128 quads = the 128 TMEM lanes, N = 4*CIN (channel, parity) columns; K = 4*COUT runs in stages of 8 output
//   channels (32 k = 4 MMA k-steps).  Producer thread (quad m, half h) loads the 2x2 neighbourhoods of 4 channels,
//   splits them and writes 16 hi + 16 lo values into its TMEM lane (A operand in tensor memory); the prepared
//   weights of all stages stay in shared memory.  Per k-step:  a_lo x w_hi,  a_hi x w_lo,  a_hi x w_hi  (N = 4*CIN).
//   Two accumulators alternate, so the epilogue of a tile (LeakyReLU mask, 8-byte stores) overlaps the MMAs of the next.
//   TMEM columns: accumulators [0, 4*CIN) and [4*CIN, 8*CIN), then two A buffers of [32 hi | 32 lo].
// -------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
constexpr size_t conv_tc_dgrad_smem() { return (size_t)((COUT / 8) * 8 * 2 * 4 * CIN * 4) * sizeof(float); }

// the data-gradient weights (prepared in the forward pass of the same step: final long before any backward kernel starts) into
// shared memory; called ahead of griddepcontrol.wait
template <int CIN, int COUT>
__device__ __forceinline__ void conv_tc_dgrad_preload(const float* __restrict__ wprep, float* smem) {
    constexpr int NST = COUT / 8, N = 4 * CIN, B_STAGE = 8 * 2 * N * 4;
    const int t = threadIdx.x;
    if (t < 256) {
        for (int i = t; i < NST * B_STAGE / 4; i += 256) cp_async16(smem + 4 * i, wprep + 4 * i);
        cp_async_commit();
    }
}
template <int CIN, int COUT, int HIN>
__device__ __forceinline__ void conv_tc_dgrad_body(const int cta, const int nctas, const float* __restrict__ dpre,
                                                   const float* __restrict__ wprep, const float* __restrict__ act_in,
                                                   float* __restrict__ din, int total_q, float* smem, BwdCtl* ctl) {
    // (the prepared weights were requested by conv_tc_bwd ahead of the grid dependency: conv_tc_dgrad_preload)
    constexpr int HO = HIN / 2, NST = COUT / 8, N = 4 * CIN;
    constexpr int B_STAGE = 8 * 2 * N * 4;  // floats: [8 slots][hi rows | lo rows][4]
    constexpr uint32_t A_COL0 = 2 * N, A_BUF = 64, A_LO = 32;
    static_assert(NST % 2 == 0 && A_COL0 + 2 * A_BUF <= kBwdTmemCols, "layout");
    float* sB = smem;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t tmem = ctl->tmem_base;
    const int ntiles = (total_q + 127) / 128;
    const int my_tiles = (cta < ntiles) ? (ntiles - 1 - cta) / nctas + 1 : 0;

    if (warp < 8) {
        const int m = t & 127, h = t >> 7, q = warp & 3, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;

        struct Raw { float v[16]; };  // [channel e][nb = da*2 + dc]
        auto gather = [&](Raw& r, int tile, int c) {
            const int gq = tile * 128 + m;
            const bool valid = gq < total_q;
            const int b = gq / (HO * HO), rem = gq % (HO * HO), a = rem / HO, cq = rem % HO;
            const bool a1 = valid && (a + 1 < HO), c1 = cq + 1 < HO;
            const float* p0 = dpre + ((size_t)b * COUT + 8 * c + 4 * h) * HO * HO + rem;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float* p = p0 + (size_t)e * HO * HO;
                r.v[4 * e + 0] = valid ? __ldg(p) : 0.f;
                r.v[4 * e + 1] = (valid && c1) ? __ldg(p + 1) : 0.f;
                r.v[4 * e + 2] = a1 ? __ldg(p + HO) : 0.f;
                r.v[4 * e + 3] = (a1 && c1) ? __ldg(p + HO + 1) : 0.f;
            }
        };
        auto gather_ahead = [&](Raw& r, int it, int c) {
            c += 2;
            if (c >= NST) { c -= NST; ++it; }
            if (it < my_tiles) gather(r, cta + it * nctas, c);
        };
        int g = 0;
        auto stage = [&](Raw& r, int it, int c) {
            const int s = g & 1;
            if (g >= 2) mbar_wait(&ctl->empty[s], (uint32_t)((g >> 1) - 1) & 1u);
            fence_after_sync();
            if (g < 8) DD_TC_STAMP_C(2 * g, 0, cta);
            float hi[16], lo[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                hi[k] = tf32_rna(r.v[k]);
                lo[k] = tf32_rna(r.v[k] - hi[k]);
            }
            const uint32_t a_col = tmem + lane_base + A_COL0 + (uint32_t)s * A_BUF + (uint32_t)h * 16u;
            tmem_st16(a_col, hi);
            tmem_st16(a_col + A_LO, lo);
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(&ctl->full[s]);
            if (g < 8) DD_TC_STAMP_C(2 * g + 1, 0, cta);
            gather_ahead(r, it, c);
            ++g;
        };
        // epilogue of tile `it` (accumulator buffer it & 1): runs while the MMA warp already works on tile it + 1.  This
        // thread: columns [half*N/2, (half+1)*N/2) = CIN/2 channels x 4 parities.  The activations that gate the
        // LeakyReLU derivative do not depend on the MMAs: they are requested one tile of stages earlier (load_act).
        constexpr int NCH = CIN / 2;
        struct Act { float2 top[NCH], bot[NCH]; };
        auto out_base = [&](int it, bool& ok) {
            const int gqe = (cta + it * nctas) * 128 + q * 32 + lane;
            ok = gqe < total_q;
            const int be = gqe / (HO * HO), reme = gqe % (HO * HO), ae = reme / HO, ce = reme % HO;
            return (((size_t)be * CIN + half * NCH) * HIN + 2 * ae) * HIN + 2 * ce;
        };
        auto load_act = [&](Act& a, int it) {
            bool ok;
            const size_t base0 = out_base(it, ok);
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                a.top[i] = ok ? __ldg(reinterpret_cast<const float2*>(act_in + base0 + (size_t)i * HIN * HIN)) : make_float2(0.f, 0.f);
                a.bot[i] = ok ? __ldg(reinterpret_cast<const float2*>(act_in + base0 + (size_t)i * HIN * HIN + HIN)) : make_float2(0.f, 0.f);
            }
        };
        auto epilogue = [&](const Act& a, int it) {
            const int ab = it & 1;
            bool ok;
            const size_t base0 = out_base(it, ok);
            mbar_wait(&ctl->done[ab], (uint32_t)(it >> 1) & 1u);
            fence_after_sync();
            if (it < 2) DD_TC_STAMP_C(33 + 3 * it, 0, cta);
#pragma unroll
            for (int gg = 0; gg < N / 32; ++gg) {
                float ra[16];
                tmem_ld16(tmem + lane_base + (uint32_t)(ab * N + half * (N / 2) + gg * 16), ra);
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int ch = gg * 4 + i;
                        const size_t base = base0 + (size_t)ch * HIN * HIN;
                        float2 top = make_float2(leaky_grad(a.top[ch].x, ra[4 * i + 0]), leaky_grad(a.top[ch].y, ra[4 * i + 1]));
                        float2 bot = make_float2(leaky_grad(a.bot[ch].x, ra[4 * i + 2]), leaky_grad(a.bot[ch].y, ra[4 * i + 3]));
                        *reinterpret_cast<float2*>(din + base) = top;
                        *reinterpret_cast<float2*>(din + base + HIN) = bot;
                    }
                }
            }
            fence_before_sync();
            mbar_arrive(&ctl->acc_empty[ab]);
            if (it < 2) DD_TC_STAMP_C(34 + 3 * it, 0, cta);
        };
        Raw r0, r1;
        Act act;
        if (my_tiles > 0) {
            gather(r0, cta, 0);
            gather(r1, cta, 1);
        }
        cp_async_wait_all();
        fence_proxy_async();
        for (int it = 0; it < my_tiles; ++it) {
            if (it > 0) load_act(act, it - 1);
#pragma unroll 1
            for (int c = 0; c < NST; c += 2) {
                stage(r0, it, c);
                stage(r1, it, c + 1);
            }
            if (it > 0) epilogue(act, it - 1);
        }
        if (my_tiles > 0) {
            load_act(act, my_tiles - 1);
            epilogue(act, my_tiles - 1);
        }
    } else {
        constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
        const uint32_t b_base = smem_u32(sB);
        for (int it = 0, g = 0; it < my_tiles; ++it) {
            const uint32_t ab = (uint32_t)it & 1u;
            if (it >= 2) mbar_wait(&ctl->acc_empty[ab], (uint32_t)((it >> 1) - 1) & 1u);  // the epilogue of tile it-2 has read D[ab]
            for (int c = 0; c < NST; ++c, ++g) {
                const int s = g & 1;
                mbar_wait(&ctl->full[s], (uint32_t)(g >> 1) & 1u);
                fence_after_sync();
                if (g < 8) DD_TC_STAMP_C(16 + 2 * g, 256, cta);
                if (elect_one()) {
                    const uint32_t a0 = tmem + A_COL0 + (uint32_t)s * A_BUF, b0 = b_base + (uint32_t)c * (B_STAGE * 4u);
                    const uint32_t d = tmem + ab * N;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t bo = b0 + (uint32_t)(2 * j) * (2 * N * 16u);
                        const uint64_t b_hi = make_desc(bo, 2 * N * 16u, 128u), b_lo = make_desc(bo + N * 16u, 2 * N * 16u, 128u);
                        mma_tf32_ts(d, a0 + A_LO + 8u * j, b_hi, idesc, (c > 0 || j > 0) ? 1u : 0u);  // lo x hi
                        mma_tf32_ts(d, a0 + 8u * j, b_lo, idesc, 1u);                                 // hi x lo
                        mma_tf32_ts(d, a0 + 8u * j, b_hi, idesc, 1u);                                 // hi x hi
                    }
                    mma_commit(&ctl->empty[s]);
                    if (c == NST - 1) mma_commit(&ctl->done[ab]);
                }
                if (g < 8) DD_TC_STAMP_C(17 + 2 * g, 256, cta);
                __syncwarp();
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// weight + bias gradient.  Rows r = tap*CIN + ci (r = 9*CIN: the row of ones -> bias gradient), padded to RP.
//   CTA = a contiguous range of pixel tiles (PXT pixels = PXT/8 MMA k-steps each); accumulators for all rows stay in
//   TMEM (NBLK blocks of 128 rows; the last block overlaps its predecessor so that every block is a full M = 128)
//   until the range is done, then one slice [RP][COUT] of partial sums goes to global memory.
//   Producer thread (pixel p = t % PXT, group w = t / PXT): 4 channels x 9 taps.  Both operands are K-major with
//   K = pixels: [chunk = pixel/4][row][4 pixels], 4-byte stores; the row pitch is 1 (mod 8) chunks so that the 32
//   lanes of a warp (consecutive pixels, one row) hit 32 different banks.  Two operand buffers alternate.
//   Per (block, k-step):  a_hi x [d_hi | d_lo] (N = 64)  and  a_lo x d_hi (N = 32).  TMEM: block b at columns 64*b.
// -------------------------------------------------------------------------------------------------------------
template <int CIN>
struct WgradCfg {
    static constexpr int PXT = 256 / (CIN / 4);              // CIN=32: 32 pixels, CIN=16: 64 pixels
    static constexpr int ROWS = 9 * CIN + 1;                 // + ones row
    static constexpr int RP = (ROWS + 3) / 4 * 4;
    static constexpr int NBLK = (RP + 127) / 128;
    static constexpr int RPAD = (RP + 7) / 8 * 8 + 1;        // smem row pitch (in 16-byte chunks), == 1 (mod 8)
    __host__ __device__ static constexpr int row0(int blk) { return blk + 1 < NBLK ? blk * 128 : RP - 128; }
};
template <int CIN, int COUT>
constexpr size_t conv_tc_wgrad_smem() {
    using Cfg = WgradCfg<CIN>;
    return (size_t)2 * (2 * (Cfg::PXT / 4) * Cfg::RPAD * 4 + (Cfg::PXT / 4) * (2 * COUT + 1) * 4) * sizeof(float);
}

template <int CIN, int COUT, int HIN>
__device__ __forceinline__ void conv_tc_wgrad_body(const int cta, const int nslices, const float* __restrict__ in,
                                                   const float* __restrict__ dpre, float* __restrict__ partial, int total_px,
                                                   float* smem, BwdCtl* ctl) {
    using Cfg = WgradCfg<CIN>;
    constexpr int HO = HIN / 2, PXT = Cfg::PXT, NBLK = Cfg::NBLK, RP = Cfg::RP, RPAD = Cfg::RPAD, CPAD = 2 * COUT + 1;
    constexpr int A_HALF = (PXT / 4) * RPAD * 4, B_SZ = (PXT / 4) * CPAD * 4;  // floats
    constexpr int BUF = 2 * A_HALF + B_SZ;
    constexpr int NDV = (COUT / 4) * PXT / 256;
    static_assert(COUT == 32 && NBLK * 64 <= (int)kBwdTmemCols, "layout");
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t tmem = ctl->tmem_base;
    const int ntiles = (total_px + PXT - 1) / PXT;
    const int t0 = (int)((long long)ntiles * cta / nslices), t1 = (int)((long long)ntiles * (cta + 1) / nslices);
    float* slice = partial + (size_t)cta * RP * COUT;

    if (warp < 8) {
        const int p = t % PXT, w = t / PXT;  // w = group of 4 input channels
        // rows 9*CIN .. RP-1 of both buffers: ones (bias gradient) + zero padding, written once
        for (int i = t; i < 2 * PXT * 4; i += 256) {
            const int bf = i / (PXT * 4), k = i % (PXT * 4), pp = k % PXT, rr = 9 * CIN + k / PXT;
            const int off = bf * BUF + ((pp >> 2) * RPAD + rr) * 4 + (pp & 3);
            smem[off] = rr == 9 * CIN ? 1.f : 0.f;
            smem[off + A_HALF] = 0.f;
        }
        struct Raw {
            float2 x[12];
            float own[12];
            float dv[NDV][4];
        };
        // `in` is an activation of the forward pass (final long before any backward kernel), dpre the gradient the predecessor
        // just wrote: the first gather of a CTA requests its `in` window, THEN waits for the grid dependency, then reads dpre
        // (ld.global.cg: ptxas moves ld.global.nc across the wait)
        bool waited = false;
        auto gather = [&](Raw& v, int tile) {
            const int gp = tile * PXT + p;
            const bool valid = gp < total_px;
            const int b = gp / (HO * HO), rem = gp % (HO * HO), oh = rem / HO, ow = rem % HO;
            const float* img = in + ((size_t)b * CIN + 4 * w) * HIN * HIN + 2 * ow;
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int ih = 2 * oh - 1 + kh;
                    const bool ok = valid && ih >= 0;
                    const float* rp = img + (size_t)e * HIN * HIN + ih * HIN;
                    v.x[e * 3 + kh] = ok ? __ldg(reinterpret_cast<const float2*>(rp)) : make_float2(0.f, 0.f);
                    v.own[e * 3 + kh] = (ok && lane == 0 && ow > 0) ? __ldg(rp - 1) : 0.f;
                }
            if (!waited) {
                pdl_wait();
                waited = true;
            }
#pragma unroll
            for (int u = 0; u < NDV; ++u) {
                const int idx = t + u * 256, pp = idx % PXT, gch = idx / PXT;
                const int gpp = tile * PXT + pp;
                const int bb = gpp / (HO * HO), rr = gpp % (HO * HO);
                const float* d = dpre + ((size_t)bb * COUT + 4 * gch) * HO * HO + rr;
#pragma unroll
                for (int e = 0; e < 4; ++e) v.dv[u][e] = gpp < total_px ? __ldcg(d + (size_t)e * HO * HO) : 0.f;
            }
        };
        int g = 0;
        auto stage = [&](Raw& v, int tile) {
            const int s = g & 1;
            if (g >= 2) mbar_wait(&ctl->empty[s], (uint32_t)((g >> 1) - 1) & 1u);
            float* sA = smem + s * BUF;
            float* sB = sA + 2 * A_HALF;
            const bool first_col = (p % HO) == 0;  // ow == 0 (HO divides PXT or PXT divides HO: the column is tile-independent)
            {
                float* dst = sA + ((p >> 2) * RPAD + 4 * w) * 4 + (p & 3);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float nb = __shfl_up_sync(0xffffffffu, v.x[e * 3 + kh].y, 1);
                        const float left = first_col ? 0.f : (lane == 0 ? v.own[e * 3 + kh] : nb);
                        const float val[3] = {left, v.x[e * 3 + kh].x, v.x[e * 3 + kh].y};
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float hi = tf32_rna(val[kw]);
                            dst[((kh * 3 + kw) * CIN + e) * 4] = hi;
                            dst[A_HALF + ((kh * 3 + kw) * CIN + e) * 4] = tf32_rna(val[kw] - hi);
                        }
                    }
            }
#pragma unroll
            for (int u = 0; u < NDV; ++u) {
                const int idx = t + u * 256, pp = idx % PXT, gch = idx / PXT;
                float* dst = sB + ((pp >> 2) * CPAD + 4 * gch) * 4 + (pp & 3);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float hi = tf32_rna(v.dv[u][e]);
                    dst[e * 4] = hi;
                    dst[(COUT + e) * 4] = tf32_rna(v.dv[u][e] - hi);
                }
            }
            fence_proxy_async();
            mbar_arrive(&ctl->full[s]);
            if (tile + 2 < t1) gather(v, tile + 2);
            ++g;
        };
        Raw v0, v1;
        if (t0 < t1) gather(v0, t0);
        if (t0 + 1 < t1) gather(v1, t0 + 1);
        if (!waited) pdl_wait();   // a CTA without tiles still zeroes its slice below
#pragma unroll 1
        for (int tile = t0; tile < t1; tile += 2) {
            stage(v0, tile);
            if (tile + 1 < t1) stage(v1, tile + 1);
        }
        if (t0 < t1) {
            mbar_wait(&ctl->done[0], 0u);
            fence_after_sync();
            const int q = warp & 3, half = warp >> 2;
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                float ra[16], rb[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 64 + half * 16), ra);
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 64 + COUT + half * 16), rb);
                const int row = Cfg::row0(blk) + q * 32 + lane;
                if (row >= blk * 128) {  // the overlapping part of the last block was already written
                    float4* o = reinterpret_cast<float4*>(slice + (size_t)row * COUT + half * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = make_float4(ra[4 * i] + rb[4 * i], ra[4 * i + 1] + rb[4 * i + 1], ra[4 * i + 2] + rb[4 * i + 2],
                                           ra[4 * i + 3] + rb[4 * i + 3]);
                }
            }
            fence_before_sync();
        } else {
            for (int i = t; i < RP * COUT; i += 256) slice[i] = 0.f;
        }
    } else {
        constexpr uint32_t idesc64 = make_idesc(128, 2 * COUT, 0, 0), idesc32 = make_idesc(128, COUT, 0, 0);
        const uint32_t base = smem_u32(smem);
        for (int tile = t0, g = 0; tile < t1; ++tile, ++g) {
            const int s = g & 1;
            mbar_wait(&ctl->full[s], (uint32_t)(g >> 1) & 1u);
            fence_after_sync();
            if (elect_one()) {
                const uint32_t a0 = base + (uint32_t)s * (BUF * 4u), b0 = a0 + 2 * A_HALF * 4u;
#pragma unroll
                for (int blk = 0; blk < NBLK; ++blk) {
#pragma unroll
                    for (int ks = 0; ks < PXT / 8; ++ks) {
                        const uint32_t ao = a0 + (uint32_t)(2 * ks * RPAD + Cfg::row0(blk)) * 16u;
                        const uint64_t bd = make_desc(b0 + (uint32_t)(2 * ks * CPAD) * 16u, CPAD * 16u, 128u);
                        mma_tf32(tmem + (uint32_t)(blk * 64), make_desc(ao, RPAD * 16u, 128u), bd, idesc64, (g > 0 || ks > 0) ? 1u : 0u);
                        mma_tf32(tmem + (uint32_t)(blk * 64), make_desc(ao + A_HALF * 4u, RPAD * 16u, 128u), bd, idesc32, 1u);
                    }
                }
                mma_commit(&ctl->empty[s]);
                if (tile == t1 - 1) mma_commit(&ctl->done[0]);
            }
            __syncwarp();
        }
    }
}

template <int CIN, int COUT, int HIN>
constexpr size_t conv_tc_bwd_smem() {
    return conv_tc_wgrad_smem<CIN, COUT>() > conv_tc_dgrad_smem<CIN, COUT>() ? conv_tc_wgrad_smem<CIN, COUT>()
                                                                              : conv_tc_dgrad_smem<CIN, COUT>();
}
template <int CIN, int COUT, int HIN>
__global__ void __launch_bounds__(288, 1)
conv_tc_bwd(const float* __restrict__ in, const float* __restrict__ dpre, const float* __restrict__ wprep_dgrad,
            const float* __restrict__ act_in, float* __restrict__ partial, float* __restrict__ din, int n_w, int total_px) {
    extern __shared__ __align__(128) float smem_tc[];
    __shared__ BwdCtl ctl;
    const int t = threadIdx.x, warp = t >> 5;
    DD_TC_CTA_BEGIN();
    DD_TC_STAMP_C(63, 0, (int)blockIdx.x - n_w);
    if (warp == 8) tmem_alloc(&ctl.tmem_base, kBwdTmemCols);
    if (t == 0) {
        mbar_init(&ctl.full[0], 256); mbar_init(&ctl.full[1], 256);
        mbar_init(&ctl.empty[0], 1); mbar_init(&ctl.empty[1], 1);
        mbar_init(&ctl.done[0], 1); mbar_init(&ctl.done[1], 1);
        mbar_init(&ctl.acc_empty[0], 256); mbar_init(&ctl.acc_empty[1], 256);
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if ((int)blockIdx.x < n_w) {
        // weight-gradient CTAs wait inside their first gather (behind the loads of the forward activations).  They may release the
        // dependents right away: the data-gradient CTAs of the same grid do so only after the wait, and the dependent grid starts
        // when EVERY CTA has, i.e. when the predecessor of this grid has completed.
        pdl_launch();
        conv_tc_wgrad_body<CIN, COUT, HIN>(blockIdx.x, n_w, in, dpre, partial, total_px, smem_tc, &ctl);
    } else {
        conv_tc_dgrad_preload<CIN, COUT>(wprep_dgrad, smem_tc);
        pdl_wait();
        pdl_launch();   // after the TMEM allocation (dd_common.cuh) and after the wait (conv_tc_fwd explains why)
        conv_tc_dgrad_body<CIN, COUT, HIN>(blockIdx.x - n_w, gridDim.x - n_w, dpre, wprep_dgrad, act_in, din, total_px, smem_tc, &ctl);
    }
    fence_before_sync();
    __syncthreads();
    DD_TC_CTA_END();
    if (warp == 8) tmem_dealloc(ctl.tmem_base, kBwdTmemCols);
}

// sum of the weight-gradient slices of all five layers in index order (deterministic), one launch.
//   tensor-core layers: slices [nslices][RP][32] (row = tap*cin + ci, last live row = bias) -> the reference's
//   [co][ci][3][3] layout;  conv1 (cin == 0 here): slices [nslices][nw + nb] already in the reference layout.
struct ReduceJob {
    const float* partial;
    float* dw;
    float* db;
    int nslices, n, stride, cin, nw, block0;  // n live outputs per slice, stride floats between slices
    int off_w, off_b;                         // offsets of dw / db in the canonical flat gradient (peer exchange)
};
struct ReduceJobs {
    ReduceJob j[5];
};
constexpr int kReduceOut = 128;     // outputs per CTA: four per lane (one 128-bit load per slice)
constexpr int kReduceWarps = 8;     // 256-thread CTAs (32 registers): two of them fit on an SM BESIDE a conv1_wgrad CTA
// The slices of the tensor-core layers (jobs 1..4) were written by kernels that are at least the predecessor's predecessor, and
// the predecessor (conv1_wgrad_kernel) releases its dependents only after its own wait: those CTAs -- 253 of 257 -- do not wait
// for the grid dependency at all and run beside conv1_wgrad; only the four CTAs of conv1's own slices (job 0) wait for it.
// (With the peer exchange every CTA waits: its pushes must not overtake the previous exchange.)
__global__ void __launch_bounds__(kReduceWarps * 32)
wgrad_reduce_kernel(const ReduceJobs jobs, const PushCtx px) {
    pdl_launch();
    __shared__ float s_part[kReduceWarps][kReduceOut + 1];
    int ji = 0;
#pragma unroll
    for (int k = 1; k < 5; ++k)
        if ((int)blockIdx.x >= jobs.j[k].block0) ji = k;
    if (ji == 0 || px.world > 1) pdl_wait();
    const unsigned int tag = push_tag(px);
    const ReduceJob jb = jobs.j[ji];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int base = ((int)blockIdx.x - jb.block0) * kReduceOut;
    {   // warp `wid` sums slices wid, wid + 8, ... for outputs base + 4*lane .. + 3 (n and stride are multiples of 4)
        constexpr int NW = kReduceWarps;
        const int i = base + 4 * lane;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        auto add = [](float4& a, const float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; };
        if (i < jb.n) {
            const float* p0 = jb.partial + i;
            int s = wid;
            for (; s + 3 * NW < jb.nslices; s += 4 * NW) {
                const float4 v0 = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)s * jb.stride));
                const float4 v1 = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)(s + NW) * jb.stride));
                const float4 v2 = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)(s + 2 * NW) * jb.stride));
                const float4 v3 = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)(s + 3 * NW) * jb.stride));
                add(a0, v0); add(a1, v1); add(a2, v2); add(a3, v3);
            }
            for (; s < jb.nslices; s += NW) add(a0, __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)s * jb.stride)));
        }
        s_part[wid][4 * lane] = (a0.x + a1.x) + (a2.x + a3.x);
        s_part[wid][4 * lane + 1] = (a0.y + a1.y) + (a2.y + a3.y);
        s_part[wid][4 * lane + 2] = (a0.z + a1.z) + (a2.z + a3.z);
        s_part[wid][4 * lane + 3] = (a0.w + a1.w) + (a2.w + a3.w);
    }
    __syncthreads();
    const int i = base + wid * 32 + lane;  // warps 0..3 finish 32 outputs each
    if (wid < kReduceOut / 32 && i < jb.n) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < kReduceWarps; ++k) r += s_part[k][wid * 32 + lane];
        int off;  // canonical flat offset of this output
        if (jb.cin == 0) {
            if (i < jb.nw) { jb.dw[i] = r; off = jb.off_w + i; } else { jb.db[i - jb.nw] = r; off = jb.off_b + i - jb.nw; }
        } else {
            const int row = i >> 5, co = i & 31;
            if (row == 9 * jb.cin) { jb.db[co] = r; off = jb.off_b + co; }
            else {
                const int e = (co * jb.cin + row % jb.cin) * 9 + row / jb.cin;
                jb.dw[e] = r;
                off = jb.off_w + e;
            }
        }
        if (px.world > 1) push_grad(px, tag, off, r);  // straight into every peer's slot (peer exchange, see below)
    }
}

// Final kernel of dd_predictor_bwd_allreduce.  Every gradient element was stored into slot [rank] of every peer by the
// kernel that produced it: the FC gradients (80 % of the bytes) by fc_bwd_kernel at the START of the backward, overlapped
// with the whole convolution backward, the convolution gradients by the deferred slice reduction.  This kernel
//   1. for every gradient element polls the 8-byte {value, tag} words of all ranks until they carry this exchange's tag
//      and adds them in rank order (bitwise identical on every rank), written through the 14 tensor pointers;
//   2. its last CTA closes the exchange (epoch + 1).
struct GradPtrs {
    float* p[14];  // state-dict order
};
__global__ void __launch_bounds__(512)
allreduce_exchange_kernel(const PushCtx px, const GradPtrs gp) {
    pdl_begin();
    ExchangeHeader* me = reinterpret_cast<ExchangeHeader*>(px.buf[px.rank]);
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&me->epoch), tag = e + 1u;
    constexpr int kOff[15] = {grad_off_conv_w(0), grad_off_conv_b(0), grad_off_conv_w(1), grad_off_conv_b(1), grad_off_conv_w(2),
                              grad_off_conv_b(2), grad_off_conv_w(3), grad_off_conv_b(3), grad_off_conv_w(4), grad_off_conv_b(4),
                              kGradOffFc1W, kGradOffFc1B, kGradOffFc2W, kGradOffFc2B, kNumGrads};
    auto segment = [&](int i) {
        int seg = 0;
#pragma unroll
        for (int k = 1; k < 14; ++k)
            if (i >= kOff[k]) seg = k;
        return seg;
    };
    const uint2* slots = exchange_slot(px.buf[px.rank], e & 1u, 0);
    // wait gently for the peers first: one thread per CTA watches one word that every rank pushes from its LAST producing
    // kernel (the slice reduction) and backs off between polls; without this an early rank hammers its L2 with 75 k polling
    // threads exactly while the late ranks' stores are arriving.  The per-element tag checks below stay (stores of
    // different CTAs are not ordered), but then they almost always pass at the first load.
    if (threadIdx.x == 0) {
        for (int r = 0; r < px.world; ++r) {
            const uint2* src = slots + (size_t)r * kGradPad + grad_off_conv_w(0);
            unsigned int t = 0, spin = 0;
            do {
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(t) : "l"(&src->y) : "memory");
                if (t != tag) __nanosleep(200);
                if (++spin > (1u << 26)) __trap();
            } while (t != tag);
        }
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kNumGrads; i += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < px.world; ++r) {
            const uint2* src = slots + (size_t)r * kGradPad + i;
            unsigned long long wv;  // {value, tag} read with ONE 64-bit load (single-copy atomic, see push_grad)
            unsigned int spin = 0;
            do {
                asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(wv) : "l"(src) : "memory");
                if (++spin > (1u << 27)) __trap();  // a peer never arrived (about a minute): fail the launch instead of hanging the device
            } while ((unsigned int)(wv >> 32) != tag);
            acc += __uint_as_float((unsigned int)wv);
        }
        const int seg = segment(i);
        gp.p[seg][i - kOff[seg]] = acc;
    }
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&me->blocks_done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {  // every element of every rank has been read: the exchange is complete on this rank
        me->blocks_done = 0u;
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(&me->epoch) = tag;
    }
}

}  // namespace tc
}  // namespace dd
