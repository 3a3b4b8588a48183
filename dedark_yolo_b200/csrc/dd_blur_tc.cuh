// dd_blur_tc.cuh -- the 25x25 Gaussian of the USM filter (nn/modules/filtersB.py:154-175) and its adjoint on the 5th-generation
// tensor cores: both separable 25-tap passes are banded-Toeplitz GEMMs (tcgen05.mma kind::tf32, fp32 accumulation in TMEM).
//
// One "block" = R staged rows x 152 staged columns of one (plane, 128-column strip); blocks march down a segment of rows:
//
//   pass 1 (horizontal)   D1T[c][r] = sum_c'  KhT[c][c'] * T[r][c']        c  = 128 output columns = the 128 TMEM lanes
//        A = KhT, the banded constant (tap(c' - c)), kept in TENSOR MEMORY for the whole kernel (written once per CTA)
//        B = the staged tile T in shared memory, K-major (k = c'), no-swizzle core-matrix layout [16 B column chunk][row][4]
//        D = R new columns of a ring of R + 32 TMEM columns (one column per staged row)
//   pass 2 (vertical)     OutT[c][n] = sum_r'  D1T[c][r'] * Kv[n][r']      n  = R output rows
//        A = the ring itself -- the accumulator of pass 1 is the A operand of pass 2 without leaving tensor memory
//        B = a sliding window of ONE small constant matrix G[e][kk] = tap(kk - e + CC) in shared memory; only the 32 output
//            rows inside the band of a k-step are touched (N = 32), the first k-step of a block runs at N = R to initialise
//   epilogue               thread <-> lane c reads OutT[c][n] (tcgen05.ld) -- consecutive lanes are consecutive image columns, so
//                          the stores of one output row are coalesced.
//
// Precision.  X3 = true ("3xTF32", fp32 I/O mode, 1e-5 gate): every operand is split x = hi + lo (both exact in TF32) and every
// product is hi*hi + hi*lo + lo*hi; the ring is split between the passes (tcgen05.ld -> 2 integer ops -> tcgen05.st).
// X3 = false (bf16 I/O mode, 2e-2 gate): one MMA per k-step, the ring feeds pass 2 untouched.
//
// Reflect padding (F.pad(mode='reflect'), filtersB.py:167) is materialised in the staged tile (mirrored halo columns, reflected
// halo rows), so the interior Toeplitz constants serve every tile.  The ADJOINT of the reflect-padded blur (the backward) is
// the same operator applied to the mirror-extended cotangent with the image-border row / column doubled, followed by halving
// the outputs on the border row / column:  with W(i,j) = k[j-i] + [j>=1] k[j+i] + [j<=n-2] k[2(n-1)-j-i]  (dd_recovery_bwd.cu),
// sum_i W(i,j) g[i] = sum_i' k[j-i'] g^[i'] for 1 <= j <= n-2 where g^[-i] = g[i], g^[0] = 2 g[0], and exactly twice the true
// value for j = 0 and j = n-1.
#pragma once
#include "dd_tcgen05.cuh"
#include "dd_recovery.cuh"

namespace dd {
namespace btc {

using tc::elect_one;
using tc::fence_after_sync;
using tc::fence_before_sync;
using tc::make_desc;
using tc::make_idesc;
using tc::mma_commit;
using tc::mma_tf32_ts;
using tc::tf32_rna;

constexpr int kCW = 16;                  // compute warps (stage, split, epilogue)
constexpr int kCT = kCW * 32;            // 512 compute threads
constexpr int kThreadsTC = kCT + 32;     // + the MMA warp (warp 16)
constexpr int kKC = kStripW + 2 * kRadius;  // 152 staged columns = K of pass 1 (19 k-steps)
constexpr int kCH = kKC / 4;             // 38 column chunks of 16 bytes
constexpr int kNBuf = 3;                 // tile buffers: block b+2 is staged while block b is between its two passes
constexpr int kSide = 4;                 // side slots: the last 12 staged rows of a block (centre values of the next block's first outputs)
constexpr int kSchedCtasTC = 148;        // one persistent CTA per SM (the kernel owns all 512 TMEM columns of its SM)

template <int R_, bool X3_>
struct Lay {
    static constexpr int R = R_;
    static constexpr bool X3 = X3_;
    static constexpr int RING = R + 32;  // >= R + 24, multiple of 16 so that a block wraps in pieces of N % 16 == 0
    static constexpr uint32_t A_HI = 0, A_LO = kKC, RING_HI = X3 ? 2 * kKC : kKC, RING_LO = RING_HI + RING,
                              OUT = RING_HI + (X3 ? 2 : 1) * RING, COLS_USED = OUT + R, TMEM_COLS = 512;
    static constexpr int T_LBO = R * 16 + 16;             // bytes between column chunks of a tile (+16: conflict-free column reads)
    static constexpr int T_BYTES = kCH * T_LBO;           // one tile (hi or lo)
    static constexpr int T_BUF = (X3 ? 2 : 1) * T_BYTES;  // one buffer: [hi | lo]
    static constexpr int CC = R + 16, G_ROWS = 2 * R + 16, G_LBO = G_ROWS * 16, G_BYTES = 2 * G_LBO;  // G: one of hi / lo
    static constexpr int SIDE_BYTES = kRadius * kStripW * 4;
    static constexpr int KS2 = (R + 2 * kRadius) / 8;     // k-steps of pass 2
    static constexpr int RPW = R / 4;                     // ring / output columns per warp of a lane quarter (4 warps per quarter)
    static constexpr size_t SMEM = (size_t)kNBuf * T_BUF + (size_t)kSide * SIDE_BYTES + (size_t)(X3 ? 2 : 1) * G_BYTES;
    static_assert(R % 16 == 0 && RING % 16 == 0 && COLS_USED <= 512 && RPW % 4 == 0 && R >= 32, "layout");
};

struct Ctl {
    uint64_t tile_full[kNBuf], p1_done[kNBuf], split_done, p2_done, out_empty;
    uint32_t tmem_base;
};

// ---- tcgen05.ld / st, 4 columns ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float* v) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// named barrier of the compute warps only (the MMA warp never joins)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }

__host__ __device__ __forceinline__ constexpr float tap_off(int d) {  // window position d in 0..24, else 0
    return (d >= 0 && d < kTaps) ? tapj(d) : 0.f;
}

// ---- one-time set-up (all compute threads; the MMA warp allocates tensor memory) ------------------------------------------------
// G[e][kk] = tap(kk - e + CC), K-major no-swizzle: 16-byte chunk kk/4 at chunk * G_LBO, row e at e * 16
template <class L>
__device__ __forceinline__ void build_constants(unsigned char* g_hi, unsigned char* g_lo, uint32_t tmem, int tid) {
    for (int i = tid; i < L::G_ROWS * 8; i += kCT) {
        const int e = i >> 3, kk = i & 7;
        const float v = tap_off(kk - e + L::CC);
        const float h = tf32_rna(v);
        const int off = (kk >> 2) * L::G_LBO + e * 16 + (kk & 3) * 4;
        *reinterpret_cast<float*>(g_hi + off) = h;
        if (L::X3) *reinterpret_cast<float*>(g_lo + off) = tf32_rna(v - h);
    }
    if (tid < 128) {  // warps 0..3: lane quarter = warp; KhT[c][c'] = tap(c' - c) into TMEM lane c, columns A_HI + c' (A_LO + c')
        const int c = tid;
        const uint32_t base = tmem + ((uint32_t)(tid & ~31) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < kKC; c0 += 4) {
            float hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = tap_off(c0 + i - c);
                hi[i] = tf32_rna(v);
                lo[i] = tf32_rna(v - hi[i]);
            }
            tc::tmem_st4(base + L::A_HI + c0, hi);
            if (L::X3) tc::tmem_st4(base + L::A_LO + c0, lo);
        }
        tc::tmem_st_wait();
    }
}

// ---- the MMA warp: one elected thread issues both passes of block `b` (segment-local index) ------------------------------------
// tile: shared-memory address of the block's buffer ([hi | lo]); g_hi / g_lo: shared-memory addresses of G
template <class L>
__device__ __forceinline__ void issue_pass1(uint32_t tmem, uint32_t tile, int b) {
    const int start = (L::R * b) % L::RING;
    const int len1 = min(L::R, L::RING - start);
#pragma unroll 1
    for (int piece = 0; piece < 2; ++piece) {
        const int n_off = piece ? len1 : 0, n = piece ? L::R - len1 : len1, col = piece ? 0 : start;
        if (n == 0) break;
        const uint32_t idesc = make_idesc(128, n, 0, 0);
        const uint32_t d = tmem + L::RING_HI + (uint32_t)col;
#pragma unroll 1
        for (int s = 0; s < kKC / 8; ++s) {
            const uint32_t bt = tile + (uint32_t)(2 * s) * L::T_LBO + (uint32_t)n_off * 16u;
            const uint64_t b_hi = make_desc(bt, L::T_LBO, 128u);
            mma_tf32_ts(d, tmem + L::A_HI + 8u * s, b_hi, idesc, s > 0 ? 1u : 0u);
            if (L::X3) {
                mma_tf32_ts(d, tmem + L::A_HI + 8u * s, make_desc(bt + L::T_BYTES, L::T_LBO, 128u), idesc, 1u);
                mma_tf32_ts(d, tmem + L::A_LO + 8u * s, b_hi, idesc, 1u);
            }
        }
    }
}
template <class L>
__device__ __forceinline__ void issue_pass2(uint32_t tmem, uint32_t g_hi, uint32_t g_lo, int b) {
    bool first = true;
#pragma unroll 1
    for (int s = 0; s < L::KS2; ++s) {
        const int v = L::R * b - 2 * kRadius + 8 * s;  // first staged (virtual) row of this k-step
        if (v < 0) continue;                           // rows above the segment: nothing staged (first block only)
        const int ring = v % L::RING;
        const int n0 = first ? 0 : min(max(8 * s - 2 * kRadius, 0), L::R - 32);
        const int n = first ? L::R : 32;
        const uint32_t row = (uint32_t)(n0 + L::CC - 8 * s);
        const uint32_t idesc = make_idesc(128, n, 0, 0);
        const uint32_t d = tmem + L::OUT + (uint32_t)n0;
        const uint64_t b_hi = make_desc(g_hi + row * 16u, L::G_LBO, 128u);
        mma_tf32_ts(d, tmem + L::RING_HI + (uint32_t)ring, b_hi, idesc, first ? 0u : 1u);
        if (L::X3) {
            mma_tf32_ts(d, tmem + L::RING_HI + (uint32_t)ring, make_desc(g_lo + row * 16u, L::G_LBO, 128u), idesc, 1u);
            mma_tf32_ts(d, tmem + L::RING_LO + (uint32_t)ring, b_hi, idesc, 1u);
        }
        first = false;
    }
}

// the MMA warp's share of one segment of nB blocks (gb = running block counter of the CTA, advanced here)
template <class L>
__device__ __forceinline__ void mma_warp_segment(Ctl& ctl, uint32_t tmem, unsigned char* tiles, uint32_t g_hi, uint32_t g_lo, int nB,
                                                 uint32_t& gb) {
    for (int b = 0; b < nB; ++b, ++gb) {
        const uint32_t buf = gb % kNBuf, par3 = (gb / kNBuf) & 1u;
        mbar_wait(&ctl.tile_full[buf], par3);
        fence_after_sync();
        if (elect_one()) {
            issue_pass1<L>(tmem, smem_u32(tiles + (size_t)buf * L::T_BUF), b);
            mma_commit(&ctl.p1_done[buf]);
        }
        __syncwarp();
        if (L::X3) mbar_wait(&ctl.split_done, gb & 1u);       // the ring columns of this block are split into hi / lo
        if (gb > 0) mbar_wait(&ctl.out_empty, (gb - 1) & 1u);  // the previous block's accumulator has been read
        fence_after_sync();
        if (elect_one()) {
            issue_pass2<L>(tmem, g_hi, g_lo, b);
            mma_commit(&ctl.p2_done);
        }
        __syncwarp();
    }
}

// set-up shared by the forward and the backward kernel: barriers, tensor memory, constants.  Returns the TMEM base.
template <class L>
__device__ __forceinline__ uint32_t setup(Ctl& ctl, unsigned char* g_hi, unsigned char* g_lo) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == kCW) tc::tmem_alloc(&ctl.tmem_base, L::TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < kNBuf; ++i) {
            mbar_init(&ctl.tile_full[i], kCT);
            mbar_init(&ctl.p1_done[i], 1);
        }
        mbar_init(&ctl.split_done, kCT);
        mbar_init(&ctl.p2_done, 1);
        mbar_init(&ctl.out_empty, kCT);
        fence_mbar_init();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = ctl.tmem_base;
    if (warp < kCW) {
        build_constants<L>(g_hi, g_lo, tmem, tid);
        fence_proxy_async();
        fence_before_sync();
    }
    __syncthreads();
    fence_after_sync();
    return tmem;
}
template <class L>
__device__ __forceinline__ void teardown(uint32_t tmem) {
    fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == kCW) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

// ---- between the passes (X3 only): split the R new ring columns into hi (in place) and lo ---------------------------------------
template <class L>
__device__ __forceinline__ void split_ring(uint32_t tmem, int b, int warp) {
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int start = (L::R * b) % L::RING + (warp >> 2) * L::RPW;
    float v[L::RPW];
#pragma unroll
    for (int i = 0; i < L::RPW; i += 4) tmem_ld4_nowait(lane_base + L::RING_HI + (uint32_t)((start + i) % L::RING), v + i);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < L::RPW; i += 4) {
        float hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hi[k] = tf32_rna(v[i + k]);
            lo[k] = tf32_rna(v[i + k] - hi[k]);
        }
        const uint32_t col = (uint32_t)((start + i) % L::RING);
        tc::tmem_st4(lane_base + L::RING_HI + col, hi);
        tc::tmem_st4(lane_base + L::RING_LO + col, lo);
    }
    tc::tmem_st_wait();
}

// ---- staged tile addressing ---------------------------------------------------------------------------------------------------
// element (staged row r, staged column sc) of a tile: chunk sc/4 at chunk * T_LBO, row at r * 16, (sc % 4) * 4
template <class L>
__device__ __forceinline__ int tile_off(int r, int sc) { return (sc >> 2) * L::T_LBO + r * 16 + (sc & 3) * 4; }

// store four consecutive staged columns (one chunk) of one row, split when X3
template <class L>
__device__ __forceinline__ void tile_store4(unsigned char* tile, int r, int chunk, const float* o) {
    unsigned char* p = tile + chunk * L::T_LBO + r * 16;
    if (L::X3) {
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = tf32_rna(o[i]);
        *reinterpret_cast<float4*>(p) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(p + L::T_BYTES) = make_float4(tf32_rna(o[0] - h[0]), tf32_rna(o[1] - h[1]), tf32_rna(o[2] - h[2]), tf32_rna(o[3] - h[3]));
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
}
template <class L>
__device__ __forceinline__ void tile_store1(unsigned char* tile, int r, int sc, float v) {
    unsigned char* p = tile + tile_off<L>(r, sc);
    if (L::X3) {
        const float h = tf32_rna(v);
        *reinterpret_cast<float*>(p) = h;
        *reinterpret_cast<float*>(p + L::T_BYTES) = tf32_rna(v - h);
    } else {
        *reinterpret_cast<float*>(p) = v;
    }
}
template <class L>
__device__ __forceinline__ float tile_load1(const unsigned char* tile, int r, int sc) {
    const unsigned char* p = tile + tile_off<L>(r, sc);
    float v = *reinterpret_cast<const float*>(p);
    if (L::X3) v += *reinterpret_cast<const float*>(p + L::T_BYTES);
    return v;
}

// Reflect halo of one staged row: the owner of the in-image columns col .. col+3 mirrors them into the staged columns of the
// image columns -col (strip 0) and L + (L - col) (the strip that holds the right border); `vals` are the values to mirror.
template <class L>
__device__ __forceinline__ void mirror_cols(unsigned char* tile, int r, int gc, int c0, int Lc, const float* vals) {
    if (gc <= kRadius && c0 == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = gc + i;
            if (col >= 1 && col <= kRadius) tile_store1<L>(tile, r, kRadius - col, vals[i]);
        }
    }
    if (gc + 3 >= Lc - kRadius) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = Lc - (gc + i);
            const int t = Lc + d - c0 + kRadius;
            if (d >= 1 && d <= kRadius && t < kKC) tile_store1<L>(tile, r, t, vals[i]);
        }
    }
}
// does the staged chunk that starts at image column gc (outside [0, W)) receive mirror writes?
__device__ __forceinline__ bool chunk_is_mirrored(int gc, int W) { return (gc < 0 && gc >= -kRadius) || (gc >= W && gc < W + kRadius); }

}  // namespace btc
}  // namespace dd
