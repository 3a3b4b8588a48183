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
// value for j = 0 and j = n-1 -- provided the two borders are more than one radius apart (n >= 14; n = 13, the smallest size
// reflect padding accepts, is routed to the CUDA-core kernel).
#pragma once
#include <cuda_bf16.h>

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

// optional in-kernel time stamps (profiles/microbench/blur_tc_probe.cu defines DD_BTC_TIMING); no code otherwise.
// Layout: [0] kernel start, [1] set-up done, [2] kernel end; block i of CTA 0 (i < 24):
//   compute warp 0 lane 0: 16 + 8 i + {0 loop top, 1 pass 1 awaited, 2 split done, 3 pass 2 awaited, 4 epilogue done, 5 barrier, 6 stage done}
//   MMA warp lane 0:      256 + 4 i + {0 tile awaited, 1 pass 1 issued, 2 split / accumulator awaited, 3 pass 2 issued}
#ifdef DD_BTC_TIMING
#ifndef DD_BTC_CTA
#define DD_BTC_CTA 0
#endif
__device__ long long g_btc_stamp[512];
#define BTC_STAMP(slot)                                                                          \
    do {                                                                                        \
        if (blockIdx.x == DD_BTC_CTA && (threadIdx.x & 31) == 0 && (slot) >= 0 && (slot) < 512) g_btc_stamp[(slot)] = clock64(); \
    } while (0)
#else
#define BTC_STAMP(slot) \
    do {                \
    } while (0)
#endif

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
    static constexpr int S_LBO = kRadius * 16 + 16;       // side slot: [32 centre chunks][12 rows][4 floats], same row stride as a tile
    static constexpr int SIDE_BYTES = (kStripW / 4) * S_LBO;
    static constexpr int KS2 = (R + 2 * kRadius) / 8;     // k-steps of pass 2
    static constexpr int RPW = R / 4;                     // ring / output columns per warp of a lane quarter (4 warps per quarter)
    static constexpr size_t SMEM = (size_t)kNBuf * T_BUF + (size_t)kSide * SIDE_BYTES + (size_t)(X3 ? 2 : 1) * G_BYTES;
    static_assert(R % 16 == 0 && RING % 16 == 0 && COLS_USED <= 512 && RPW % 4 == 0 && R >= 32, "layout");
};

struct Ctl {
    uint64_t tile_full[kNBuf], p1_done[kNBuf], split_done, p2_done, out_empty;
    uint32_t tmem_base;
    float tap[32];  // tap[d] = k1[|d - 12|] for window positions d = 0..24, 0 beyond (set-up only)
};

static __constant__ float c_tap25[kTaps] = {DD_K12, DD_K11, DD_K10, DD_K9, DD_K8, DD_K7, DD_K6, DD_K5, DD_K4, DD_K3, DD_K2, DD_K1, DD_K0,
                                      DD_K1,  DD_K2,  DD_K3,  DD_K4, DD_K5, DD_K6, DD_K7, DD_K8, DD_K9, DD_K10, DD_K11, DD_K12};

// ---- tcgen05.ld / st, 4 columns ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, float* v) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mbarrier traffic per WARP, not per thread: 512 arrivals on one barrier word are 512 serialised shared-memory atomics (measured:
// ~2000 cycles per barrier), and 512 pollers contend for it.  Lane 0 arrives after __syncwarp() has ordered the warp's prior
// writes (each thread has already executed the proxy / tcgen05 fence its own writes need); lane 0 polls and __syncwarp()
// releases the other lanes.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// named barrier of the compute warps only (the MMA warp never joins)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }

// ---- one-time set-up (all compute threads; the MMA warp allocates tensor memory) ------------------------------------------------
// G[e][kk] = tap(kk - e + CC), K-major no-swizzle: 16-byte chunk kk/4 at chunk * G_LBO, row e at e * 16
template <class L>
__device__ __forceinline__ void build_constants(const float* tap, unsigned char* g_hi, unsigned char* g_lo, uint32_t tmem, int tid) {
    auto tap_off = [&](int d) { return (d >= 0 && d < kTaps) ? tap[d] : 0.f; };
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
#pragma unroll 2
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
    constexpr uint64_t kStep = (uint64_t)((2 * L::T_LBO) >> 4), kLo = (uint64_t)(L::T_BYTES >> 4);  // descriptor address increments
#pragma unroll 1
    for (int piece = 0; piece < 2; ++piece) {
        const int n_off = piece ? len1 : 0, n = piece ? L::R - len1 : len1, col = piece ? 0 : start;
        if (n == 0) break;
        const uint32_t idesc = make_idesc(128, n, 0, 0);
        const uint32_t d = tmem + L::RING_HI + (uint32_t)col;
        const uint64_t d0 = make_desc(tile + (uint32_t)n_off * 16u, L::T_LBO, 128u);
#pragma unroll
        for (int s = 0; s < kKC / 8; ++s) {
            const uint64_t b_hi = d0 + (uint64_t)s * kStep;
            mma_tf32_ts(d, tmem + L::A_HI + 8u * s, b_hi, idesc, s > 0 ? 1u : 0u);
            if (L::X3) {
                mma_tf32_ts(d, tmem + L::A_HI + 8u * s, b_hi + kLo, idesc, 1u);
                mma_tf32_ts(d, tmem + L::A_LO + 8u * s, b_hi, idesc, 1u);
            }
        }
    }
}
template <class L>
__device__ __forceinline__ void issue_pass2(uint32_t tmem, uint32_t g_hi, uint32_t g_lo, int b) {
    // ring column of the first staged row of k-step 0 (v = R b - 24; negative only for b = 0, whose first three k-steps are skipped)
    const int v0 = L::R * b - 2 * kRadius;
    const int ring0 = ((v0 % L::RING) + L::RING) % L::RING;
    const uint64_t dg_hi = make_desc(g_hi, L::G_LBO, 128u), dg_lo = make_desc(g_lo, L::G_LBO, 128u);
    constexpr uint32_t idesc_full = make_idesc(128, L::R, 0, 0), idesc_band = make_idesc(128, 32, 0, 0);
    const int s_first = b == 0 ? 2 * kRadius / 8 : 0;
#pragma unroll
    for (int s = 0; s < L::KS2; ++s) {
        if (s < s_first) continue;  // rows above the segment: nothing staged (first block only)
        const bool first = s == s_first;
        int ring = ring0 + 8 * s;
        ring = ring >= L::RING ? ring - L::RING : ring;
        const int n0_band = (8 * s - 2 * kRadius) < 0 ? 0 : ((8 * s - 2 * kRadius) > L::R - 32 ? L::R - 32 : 8 * s - 2 * kRadius);
        const int n0 = first ? 0 : n0_band;
        const uint64_t row = (uint64_t)(n0 + L::CC - 8 * s);  // 16-byte units: one row of G per unit
        const uint32_t idesc = first ? idesc_full : idesc_band;
        const uint32_t d = tmem + L::OUT + (uint32_t)n0;
        mma_tf32_ts(d, tmem + L::RING_HI + (uint32_t)ring, dg_hi + row, idesc, first ? 0u : 1u);
        if (L::X3) {
            mma_tf32_ts(d, tmem + L::RING_HI + (uint32_t)ring, dg_lo + row, idesc, 1u);
            mma_tf32_ts(d, tmem + L::RING_LO + (uint32_t)ring, dg_hi + row, idesc, 1u);
        }
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
        BTC_STAMP(gb < 24 ? 256 + 4 * (int)gb : -1);
        if (elect_one()) {
            issue_pass1<L>(tmem, smem_u32(tiles + (size_t)buf * L::T_BUF), b);
            mma_commit(&ctl.p1_done[buf]);
        }
        __syncwarp();
        BTC_STAMP(gb < 24 ? 257 + 4 * (int)gb : -1);
        if (L::X3) mbar_wait(&ctl.split_done, gb & 1u);       // the ring columns of this block are split into hi / lo
        if (gb > 0) mbar_wait(&ctl.out_empty, (gb - 1) & 1u);  // the previous block's accumulator has been read
        fence_after_sync();
        BTC_STAMP(gb < 24 ? 258 + 4 * (int)gb : -1);
        if (elect_one()) {
            issue_pass2<L>(tmem, g_hi, g_lo, b);
            mma_commit(&ctl.p2_done);
        }
        __syncwarp();
        BTC_STAMP(gb < 24 ? 259 + 4 * (int)gb : -1);
    }
}

// set-up shared by the forward and the backward kernel: barriers, tensor memory, constants.  Returns the TMEM base.
template <class L>
__device__ __forceinline__ uint32_t setup(Ctl& ctl, unsigned char* g_hi, unsigned char* g_lo) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) BTC_STAMP(0);
    if (warp == kCW) tc::tmem_alloc(&ctl.tmem_base, L::TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < kNBuf; ++i) {
            mbar_init(&ctl.tile_full[i], kCW);
            mbar_init(&ctl.p1_done[i], 1);
        }
        mbar_init(&ctl.split_done, kCW);
        mbar_init(&ctl.p2_done, 1);
        mbar_init(&ctl.out_empty, kCW);
        fence_mbar_init();
    }
    if (tid < 32) ctl.tap[tid] = tid < kTaps ? c_tap25[tid] : 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = ctl.tmem_base;
    if (warp < kCW) {
        build_constants<L>(ctl.tap, g_hi, g_lo, tmem, tid);
        fence_proxy_async();
        fence_before_sync();
    }
    __syncthreads();
    fence_after_sync();
    if (tid == 0) BTC_STAMP(1);
    return tmem;
}
template <class L>
__device__ __forceinline__ void teardown(uint32_t tmem) {
    fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) BTC_STAMP(2);
    if ((threadIdx.x >> 5) == kCW) tc::tmem_dealloc(tmem, L::TMEM_COLS);
}

// ---- between the passes (X3 only): split the R new ring columns into hi (in place) and lo ---------------------------------------
// hi = x rounded to TF32 (two integer ops), lo = x - hi exactly (at most 13 significant bits; the tensor core ignores the bits
// of an operand below TF32 precision, an error of 2^-11 |lo| <= 2^-23 |x| -- no second rounding).  Groups of four columns never
// straddle the end of the ring (start and RING are multiples of 4).
template <class L>
__device__ __forceinline__ void split_ring(uint32_t tmem, int b, int warp) {
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    int col = (L::R * b) % L::RING + (warp >> 2) * L::RPW;
    col = col >= L::RING ? col - L::RING : col;
    uint32_t c[L::RPW / 4];
    float v[L::RPW];
#pragma unroll
    for (int i = 0; i < L::RPW / 4; ++i) {
        c[i] = (uint32_t)col;
        tmem_ld4_nowait(lane_base + L::RING_HI + c[i], v + 4 * i);
        col += 4;
        col = col >= L::RING ? col - L::RING : col;
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < L::RPW / 4; ++i) {
        float hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hi[k] = tf32_rna(v[4 * i + k]);
            lo[k] = v[4 * i + k] - hi[k];
        }
        tc::tmem_st4(lane_base + L::RING_HI + c[i], hi);
        tc::tmem_st4(lane_base + L::RING_LO + c[i], lo);
    }
    tc::tmem_st_wait();
}

// ---- staged tile addressing ---------------------------------------------------------------------------------------------------
// element (staged row r, staged column sc) of a tile: chunk sc/4 at chunk * T_LBO, row at r * 16, (sc % 4) * 4
template <class L>
__device__ __forceinline__ int tile_off(int r, int sc) { return (sc >> 2) * L::T_LBO + r * 16 + (sc & 3) * 4; }

// store four consecutive staged columns (one chunk) of one row, split when X3; p = address of the chunk's row in the hi tile
template <class L>
__device__ __forceinline__ void tile_store4(unsigned char* p, const float* o) {
    if (L::X3) {
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = tf32_rna(o[i]);
        *reinterpret_cast<float4*>(p) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(p + L::T_BYTES) = make_float4(o[0] - h[0], o[1] - h[1], o[2] - h[2], o[3] - h[3]);  // exact; see split_ring
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

// ---- image element types: fp32 (the reference's dtype) and bf16 (the bf16 I/O mode) -------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static __device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ float load1(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void store1_streaming(float* p, float v) { __stcs(p, v); }
    static __device__ __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float bits(unsigned short u) { return __uint_as_float((unsigned)u << 16); }
    static __device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {  // four consecutive elements, 8-byte aligned
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
    }
    static __device__ __forceinline__ float load1(const __nv_bfloat16* p) { return bits(__ldg(reinterpret_cast<const unsigned short*>(p))); }
    static __device__ __forceinline__ void store1_streaming(__nv_bfloat16* p, float v) {
        __stcs(reinterpret_cast<unsigned short*>(p), __bfloat16_as_ushort(__float2bfloat16_rn(v)));
    }
    static __device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---- the CTA's work as one stream of blocks ----------------------------------------------------------------------------------------
// A CTA owns a contiguous range of 32-row scheduling units (dd_layout.cuh) = a sequence of segments (runs of rows inside one
// plane-strip), each cut into blocks of R staged rows.  The stage, the split and the epilogue of the software pipeline walk that
// sequence with one iterator each (the stage runs three blocks ahead of the epilogue, possibly in a later segment), so the
// tensor pipe is never drained at a segment boundary.  All fields are CTA-uniform.
template <int R>
struct SegIter {
    long long blk, blk_end;
    int ord;      // ordinal of the current segment (-1 before the first block)
    int nB, b;    // blocks of the current segment, current block
    Seg u;
    __device__ __forceinline__ void init(const Sched& sc) {
        blk = sched_begin(sc, blockIdx.x);
        blk_end = sched_begin(sc, blockIdx.x + 1);
        ord = -1;
        nB = 0;
        b = 0;
    }
    // step to the next block; false at the end of the stream.  `entered`: this block is the first of a new segment.
    __device__ __forceinline__ bool next(const Sched& sc, int H, bool& entered) {
        entered = false;
        if (ord >= 0 && b + 1 < nB) {
            ++b;
            return true;
        }
        if (blk >= blk_end) return false;
        u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        nB = (u.nU + R - 1) / R;
        b = 0;
        ++ord;
        entered = true;
        return true;
    }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- staging geometry of one thread, fixed for a segment ------------------------------------------------------------------------
// The R x 38 (row, 16-byte column chunk) items of a block are dealt out as: thread t < (R/4) * 38 owns chunk j = t % 38 of the
// four rows rg + (R/4) k, k = 0..3, rg = t / 38 (R = 48: 456 of the 512 compute threads, exactly four items each, the four
// global loads of a block in flight together; the last 12 rows of a block -- the side-slot rows -- are exactly k = 3).
// Reflect padding (F.pad(mode='reflect'), filtersB.py:167): a chunk in the reflect halo loads its four mirrored source columns
// itself and then runs the same arithmetic as an in-image chunk -- no scattered mirror stores, no divergent code.
template <class L>
struct StageGeo {
    static constexpr int NG = L::R / 4;  // row groups
    static constexpr int KIND_IDLE = 0, KIND_INSIDE = 1, KIND_ZERO = 2, KIND_MIRRORED = 3;
    int rg, gc;
    int kind;        // this thread's chunk: in-image data (one 128-bit load), reflect halo (four scalar loads of the mirrored
                     // source columns), or zeros beyond the reflect range
    int toff;        // byte offset of (row rg, chunk j) inside a tile
    int soff;        // byte offset of (side row rg, chunk j) inside a side slot, -1 when the chunk is not a centre chunk
    int mc[4];       // KIND_MIRRORED: source image columns reflect(gc + i, W)
    __device__ __forceinline__ void init(int tid, int c0, int W) {
        rg = tid / kCH;
        const int j = tid - rg * kCH;
        gc = c0 - kRadius + 4 * j;
        const bool inside = gc >= 0 && gc < W;
        const bool mirrored = (gc < 0 && gc >= -kRadius) || (gc >= W && gc < W + kRadius);
        kind = tid >= NG * kCH ? KIND_IDLE : inside ? KIND_INSIDE : mirrored ? KIND_MIRRORED : KIND_ZERO;
        toff = j * L::T_LBO + rg * 16;
        soff = (j >= kRadius / 4 && j < kRadius / 4 + kStripW / 4) ? (j - kRadius / 4) * L::S_LBO + rg * 16 : -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) mc[i] = min(max(reflect(gc + i, W), 0), W - 1);
    }
    // the four staged values of (image row offset `off` = row * W, this thread's chunk); KIND_INSIDE or KIND_MIRRORED only
    template <typename T>
    __device__ __forceinline__ float4 load4(const T* __restrict__ plane, unsigned off) const {
        if (kind == KIND_INSIDE) return Elem<T>::load4(plane + off + gc);
        return make_float4(Elem<T>::load1(plane + off + mc[0]), Elem<T>::load1(plane + off + mc[1]), Elem<T>::load1(plane + off + mc[2]),
                           Elem<T>::load1(plane + off + mc[3]));
    }
};

}  // namespace btc
}  // namespace dd
