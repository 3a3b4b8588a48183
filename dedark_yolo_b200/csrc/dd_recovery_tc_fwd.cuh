// dd_recovery_tc_fwd.cuh -- a6..a12: the fused filter chain forward with the USM blur on the tensor cores (dd_blur_tc.cuh).
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,295-304,316-317.
//
// One persistent CTA per SM (16 compute warps + 1 MMA warp).  The CTAs share the marching-strip work list of the CUDA-core
// kernel (dd_layout.cuh: plane-strips cut into 32-row scheduling units, contiguous equal ranges per CTA); inside a segment the
// rows are processed in blocks of R staged rows:
//     stage(b)     x0 -> DeDark -> WB -> Gamma -> Contrast (pointwise, in registers) -> x4, split, into tile buffer b % 3
//                  (reflect halo: mirrored columns, reflected rows), last 12 rows also into a side slot
//     pass 1, [split], pass 2   on the tensor cores (MMA warp), accumulators in tensor memory
//     epilogue(b)  y = (x4 - blur) p + x4 for the R output rows whose 25-row window pass 1 has completed; thread <-> column
// software-pipelined: in iteration b the compute warps split block b, run the epilogue of block b-1 and stage block b+2 while
// the tensor pipe runs pass 2 of block b and pass 1 of block b+1.
#pragma once
#include "dd_blur_tc.cuh"

namespace dd {
namespace btc {

// DBG: identity chain (x4 = x0) and y = blur -- the bare reflect-padded 25x25 Gaussian, for the tensor-core unit test
template <int R, bool X3, bool HAS_ICA, bool FAST, bool DBG>
__global__ void __launch_bounds__(kThreadsTC, 1)
recovery_fwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                       const float* __restrict__ feat, float* __restrict__ y, int B, int H, int W) {
    using L = Lay<R, X3>;
    pdl_begin();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;
    float* side = reinterpret_cast<float*>(tiles + (size_t)kNBuf * L::T_BUF);
    unsigned char* g_hi = reinterpret_cast<unsigned char*>(side) + (size_t)kSide * L::SIDE_BYTES;
    unsigned char* g_lo = g_hi + L::G_BYTES;
    __shared__ float MS[kMaxU];  // per staged row: m = (1 - c) + c q
    __shared__ ImgParams sp;
    __shared__ Ctl ctl;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = setup<L>(ctl, g_hi, g_lo);

    const Sched sc = make_sched(B, H, W, kSchedCtasTC);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const int Lc = W - 1;
    uint32_t gb = 0;  // running block counter of this CTA (identical in every warp): barrier parities and buffer rotation

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        const int nB = (u.nU + R - 1) / R;

        if (warp == kCW) {
            mma_warp_segment<L>(ctl, tmem, tiles, smem_u32(g_hi), smem_u32(g_lo), nB, gb);
            continue;
        }

        // ================================ compute warps ================================
        compute_sync();  // the previous segment is fully drained: MS, sp, side slots and tile buffers may be rewritten
        if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
        compute_sync();
        const ChainK ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
        const float pc = sp.c, pp = sp.p;
        const float* xp = x + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
        float* yp = y + (size_t)u.plane * H * W;

        if (!DBG) {  // per-row contrast scalars of the segment (rgb2lum quirk: columns 0..2 of each row, util_filters.py:270-273)
            for (int v = tid; v < u.nU; v += kCT) {
                const int row = reflect(u.r0 - kRadius + v, H);
                float x3[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    x3[c] = chain_x3<HAS_ICA, FAST>(ck, __ldg(xp + (size_t)row * W + c), HAS_ICA ? __ldg(ip + (size_t)row * W + c) : kDefaultIcA);
                const RowLum rl = row_lum<true>(x3[0], x3[1], x3[2]);
                MS[v] = (1.f - pc) + pc * rl.q;
            }
            compute_sync();
        }

        auto stage = [&](int b) {
            const uint32_t g = gb + (uint32_t)b;
            unsigned char* tile = tiles + (size_t)(g % kNBuf) * L::T_BUF;
            float* sd = side + (size_t)(g % kSide) * (L::SIDE_BYTES / 4);
#pragma unroll 1
            for (int idx = tid; idx < R * kCH; idx += kCT) {
                const int r = idx / kCH, j = idx - r * kCH;
                const int v = R * b + r;
                const int gc = u.c0 - kRadius + 4 * j;
                const bool inside = gc >= 0 && gc < W, vrow = v < u.nU;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (inside && vrow) {
                    const int row = reflect(u.r0 - kRadius + v, H);
                    const float4 in = __ldg(reinterpret_cast<const float4*>(xp + (size_t)row * W + gc));
                    if (DBG) {
                        o[0] = in.x; o[1] = in.y; o[2] = in.z; o[3] = in.w;
                    } else {
                        float4 ic = make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                        if (HAS_ICA) ic = __ldg(reinterpret_cast<const float4*>(ip + (size_t)row * W + gc));
                        const float m = MS[v];
                        o[0] = chain_x3<HAS_ICA, FAST>(ck, in.x, ic.x) * m;
                        o[1] = chain_x3<HAS_ICA, FAST>(ck, in.y, ic.y) * m;
                        o[2] = chain_x3<HAS_ICA, FAST>(ck, in.z, ic.z) * m;
                        o[3] = chain_x3<HAS_ICA, FAST>(ck, in.w, ic.w) * m;
                    }
                    mirror_cols<L>(tile, r, gc, u.c0, Lc, o);
                }
                if (inside || !vrow || !chunk_is_mirrored(gc, W)) tile_store4<L>(tile, r, j, o);
                if (r >= R - kRadius && j >= kRadius / 4 && j < kRadius / 4 + kStripW / 4)
                    *reinterpret_cast<float4*>(sd + (r - (R - kRadius)) * kStripW + 4 * j - kRadius) = make_float4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            mbar_arrive(&ctl.tile_full[g % kNBuf]);
        };

        auto epilogue = [&](int bb) {  // outputs o = R * bb - 24 + n, n in [0, R)
            const uint32_t g = gb + (uint32_t)bb;
            mbar_wait(&ctl.p2_done, g & 1u);
            fence_after_sync();
            const int q = warp & 3, n_first = (warp >> 2) * L::RPW, c = 32 * q + lane;
            float bl[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; i += 4) tmem_ld4_nowait(tmem + ((uint32_t)(32 * q) << 16) + L::OUT + (uint32_t)(n_first + i), bl + i);
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive(&ctl.out_empty);  // the accumulator is in registers: pass 2 of the next block may overwrite it
            const unsigned char* tile = tiles + (size_t)(g % kNBuf) * L::T_BUF;
            const float* sd = side + (size_t)((g + kSide - 1) % kSide) * (L::SIDE_BYTES / 4);  // tail of block bb - 1
            const int gc = u.c0 + c;
#pragma unroll
            for (int i = 0; i < L::RPW; ++i) {
                const int o = R * bb - 2 * kRadius + n_first + i;
                if (o < 0 || o >= u.seg_len) continue;  // warp-uniform
                const int rr = o + kRadius - R * bb;     // centre row relative to block bb (negative: tail of block bb - 1)
                const float x4 = rr < 0 ? sd[(rr + kRadius) * kStripW + c] : tile_load1<L>(tile, rr, c + kRadius);
                const float yv = DBG ? bl[i] : fmaf(x4 - bl[i], pp, x4);
                if (gc < W) __stcs(yp + (size_t)(u.r0 + o) * W + gc, yv);
            }
        };

        stage(0);
        if (nB > 1) stage(1);
        for (int b = 0; b < nB; ++b) {
            const uint32_t g = gb + (uint32_t)b;
            mbar_wait(&ctl.p1_done[g % kNBuf], (g / kNBuf) & 1u);  // pass 1 of block b complete: ring columns valid, tile buffer read
            if (X3) {
                fence_after_sync();
                split_ring<L>(tmem, b, warp);
                fence_before_sync();
                mbar_arrive(&ctl.split_done);
            }
            if (b > 0) epilogue(b - 1);
            compute_sync();  // every warp is done with block b-1's buffer and side slots before they are restaged
            if (b + 2 < nB) stage(b + 2);
        }
        epilogue(nB - 1);
        gb += (uint32_t)nB;
    }

    teardown<L>(tmem);
}

template <int R, bool X3>
constexpr size_t fwd_tc_smem() { return Lay<R, X3>::SMEM; }

}  // namespace btc
}  // namespace dd
