// dd_recovery_tc_fwd.cuh -- a6..a12: the fused filter chain forward with the USM blur on the tensor cores (dd_blur_tc.cuh).
//
// Reference: nn/modules/llie.py:34-40,49-52; filtersB.py:144-259,289-303; util_filters.py:270-273,295-304,316-317.
//
// One persistent CTA per SM (16 compute warps + 1 MMA warp).  The CTAs share the marching-strip work list of the CUDA-core
// kernel (dd_layout.cuh: plane-strips cut into 32-row scheduling units, contiguous equal ranges per CTA); a CTA's range is one
// stream of blocks of R staged rows (SegIter):
//     stage(G)     x0 -> DeDark -> WB -> Gamma -> Contrast (pointwise, in registers) -> x4, split, into tile buffer G % 3
//                  (reflect halo: mirrored columns, reflected rows), last 12 rows also into a side slot
//     pass 1, [split], pass 2   on the tensor cores (MMA warp), accumulators in tensor memory
//     epilogue(G)  y = (x4 - blur) p + x4 for the R output rows whose 25-row window pass 1 has completed; thread <-> column
// software-pipelined: iteration G runs the epilogue of block G-1, issues the global loads of block G+2, splits block G (in
// the shadow of those loads) and finishes staging block G+2, while the tensor pipe runs pass 2 of block G and pass 1 of block
// G+1.  The stage may already be in the next segment (another image plane or strip): the per-segment set-up (regressors,
// per-row contrast scalars) runs inside the stream, the tensor pipe is not drained at segment boundaries.
#pragma once
#include "dd_blur_tc.cuh"

namespace dd {
namespace btc {

// DBG: identity chain (x4 = x0) and y = blur -- the bare reflect-padded 25x25 Gaussian, for the tensor-core unit test
// TX / TY: element types of x and y (float, or __nv_bfloat16 in the bf16 I/O mode; IcA, A and feat are always fp32)
template <int R, bool X3, bool HAS_ICA, bool FAST, bool DBG, typename TX = float, typename TY = float>
__global__ void __launch_bounds__(kThreadsTC, 1)
recovery_fwd_tc_kernel(const TX* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                       const float* __restrict__ feat, TY* __restrict__ y, int B, int H, int W) {
    using L = Lay<R, X3>;
    using SG = StageGeo<L>;
    static_assert(R == 48, "the stage / epilogue work split assumes 48-row blocks (12 row groups x 4 rows, 12 outputs per warp)");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;
    unsigned char* side = tiles + (size_t)kNBuf * L::T_BUF;
    unsigned char* g_hi = side + (size_t)kSide * L::SIDE_BYTES;
    unsigned char* g_lo = g_hi + L::G_BYTES;
    __shared__ float MS[kMaxU];   // per staged row of the segment being staged: m = (1 - c) + c q
    __shared__ float s_pp[4];     // USM strength p of the last four segments (the epilogue runs up to three blocks behind the stage)
    __shared__ ImgParams sp;
    __shared__ Ctl ctl;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = setup<L>(ctl, g_hi, g_lo);   // reads no global memory: overlaps the tail of the previous kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const Sched sc = make_sched(B, H, W, kSchedCtasTC);

    if (warp == kCW) {
        // ================================ MMA warp ================================
        uint32_t gb = 0;
        const long long blk_end = sched_begin(sc, blockIdx.x + 1);
        for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
            const Seg u = next_seg(blk, blk_end, sc, H);
            blk += seg_blocks(u);
            mma_warp_segment<L>(ctl, tmem, tiles, smem_u32(g_hi), smem_u32(g_lo), (u.nU + R - 1) / R, gb);
        }
    } else {
        // ================================ compute warps ================================
        SegIter<R> itS, itM, itE;   // stage, split and epilogue positions in the block stream
        itS.init(sc);
        itM.init(sc);
        itE.init(sc);
        bool entered;

        // ---- stage side -------------------------------------------------------------------------------------------------------
        // Lean by construction: a thread's four items of a block are the same chunk of rows rg, rg+12, rg+24, rg+36, so one 64-bit
        // address per block and a constant row step reach all four loads; a block whose R staged rows all lie inside the image and
        // inside the segment (all but the first and last block of an image column) takes a path without any per-row index logic.
        SG sg;
        ChainK ck;
        const TX* xp = x;
        const float* ip = IcA;
        int rowbase = 0, nU = 0;
        const size_t wstep = (size_t)SG::NG * (size_t)W;   // elements between two items of a thread
        float4 in[4], ic[4];      // the block in flight between stage_load and stage_store
        int s_vb = 0;             // its first staged row
        bool s_inner = false;     // ... and whether it takes the lean path (CTA-uniform)
        uint32_t sbuf = 0, sside = 0;   // tile buffer / side slot the next stage_store fills (block counter mod 3, mod 4)
        auto stage_enter = [&]() {  // first block of a new segment: regressors of its image, per-row contrast scalars, geometry
            const Seg& u = itS.u;
            compute_sync();  // every warp has finished staging the previous segment (MS, sp are about to be rewritten)
            if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
            compute_sync();
            ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
            const float pc = sp.c;
            if (tid == 0) s_pp[itS.ord & 3] = sp.p;
            xp = x + (size_t)u.plane * H * W;
            ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
            rowbase = u.r0 - kRadius;
            nU = u.nU;
            sg.init(tid, u.c0, W);
            if (!DBG) {  // rgb2lum quirk: columns 0..2 of each row (util_filters.py:270-273)
                for (int v = tid; v < nU; v += kCT) {
                    const size_t ro = (size_t)reflect(rowbase + v, H) * W;
                    float x3[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x3[c] = chain_x3<HAS_ICA, FAST>(ck, Elem<TX>::load1(xp + ro + c), HAS_ICA ? __ldg(ip + ro + c) : kDefaultIcA);
                    const RowLum rl = row_lum<true>(x3[0], x3[1], x3[2]);
                    MS[v] = (1.f - pc) + pc * rl.q;
                }
            }
            compute_sync();
        };
        auto stage_load = [&]() {  // the four global loads of this thread's items of block itS.b
            s_vb = R * itS.b;
            const int row0 = rowbase + s_vb;
            s_inner = row0 >= 0 && row0 + R <= H && s_vb + R <= nU;
            if (sg.kind == SG::KIND_INSIDE && s_inner) {   // one 64-bit offset per block, a constant row step between the four items
                const size_t off = (size_t)(row0 + sg.rg) * (size_t)W + (size_t)sg.gc;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    in[k] = Elem<TX>::load4(xp + off + k * wstep);
                    if (HAS_ICA) ic[k] = Elem<float>::load4(ip + off + k * wstep);
                }
                if (row0 + 2 * R <= H) {  // the rows of the next block: towards L2 while this block is processed (its loads then hit L2)
#pragma unroll
                    for (int k = 0; k < 4; ++k) prefetch_l2(xp + off + (size_t)R * (size_t)W + k * wstep);
                }
            } else if (sg.kind == SG::KIND_INSIDE || sg.kind == SG::KIND_MIRRORED) {
                const int v0 = s_vb + sg.rg;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned off = (unsigned)reflect(rowbase + min(v0 + SG::NG * k, nU - 1), H) * (unsigned)W;
                    in[k] = sg.load4(xp, off);
                    if (HAS_ICA) ic[k] = sg.load4(ip, off);
                }
            }
        };
        auto stage_item = [&](unsigned char* trow, int k, float m, bool zero) {   // pointwise chain of one item -> tile (+ side slot)
            float o[4];
            if (DBG) {
                o[0] = in[k].x; o[1] = in[k].y; o[2] = in[k].z; o[3] = in[k].w;
            } else {
                const float4 q = HAS_ICA ? ic[k] : make_float4(kDefaultIcA, kDefaultIcA, kDefaultIcA, kDefaultIcA);
                o[0] = chain_x3<HAS_ICA, FAST>(ck, in[k].x, q.x) * m;
                o[1] = chain_x3<HAS_ICA, FAST>(ck, in[k].y, q.y) * m;
                o[2] = chain_x3<HAS_ICA, FAST>(ck, in[k].z, q.z) * m;
                o[3] = chain_x3<HAS_ICA, FAST>(ck, in[k].w, q.w) * m;
            }
            if (zero) o[0] = o[1] = o[2] = o[3] = 0.f;  // rows past the segment (its last block only)
            tile_store4<L>(trow + 16 * SG::NG * k, o);
            if (k == 3 && sg.soff >= 0)  // rows R-12 .. R-1: centre values of the next block's first outputs
                *reinterpret_cast<float4*>(side + (size_t)sside * L::SIDE_BYTES + sg.soff) = make_float4(o[0], o[1], o[2], o[3]);
        };
        auto stage_store = [&]() {  // pointwise chain, split, tile buffer sbuf (+ side slot sside)
            unsigned char* trow = tiles + (size_t)sbuf * L::T_BUF + sg.toff;  // row rg; row rg + 12 k at + 192 k bytes
            if (sg.kind == SG::KIND_INSIDE || sg.kind == SG::KIND_MIRRORED) {
                const float* ms = MS + s_vb + sg.rg;
                if (s_inner) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) stage_item(trow, k, DBG ? 1.f : ms[SG::NG * k], false);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int v = s_vb + sg.rg + SG::NG * k;
                        stage_item(trow, k, DBG ? 1.f : MS[min(v, nU - 1)], v >= nU);
                    }
                }
            } else if (sg.kind == SG::KIND_ZERO) {  // halo chunks beyond the reflect range
                const float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 4; ++k) tile_store4<L>(trow + 16 * SG::NG * k, z);
            }
            fence_proxy_async();
            mbar_arrive_warp(&ctl.tile_full[sbuf]);
            sbuf = sbuf == kNBuf - 1 ? 0 : sbuf + 1;
            sside = (sside + 1) & (kSide - 1);
        };

        // ---- epilogue side ----------------------------------------------------------------------------------------------------
        const int q = warp & 3, gq = warp >> 2, c = 32 * q + lane;   // lane quarter, output-row group (12 rows), strip column
        // centre values: the first 12 outputs of a block sit on the tail of the previous block (side slot, exact fp32), the others
        // on rows 12 (gq - 1) + i of the block's own tile (hi + lo); 16 bytes between consecutive rows in both
        const int ctr_off = gq == 0 ? (c >> 2) * L::S_LBO + (c & 3) * 4 : tile_off<L>(L::RPW * (gq - 1), c + kRadius);
        const size_t ystep_b = (size_t)W * sizeof(TY);   // bytes between two output rows
        TY* ys = y;
        bool col_ok = false;
        float pp = 0.f;
        int e_r0 = 0, e_len = 0;
        uint32_t ebuf = 0, eside = kSide - 1, epar = 0;   // tile buffer / previous side slot / p2_done parity of the block in the epilogue
        int ep_ts = -100;  // time-stamp slot of the epilogue in flight (DD_BTC_TIMING only)
        auto epilogue_enter = [&]() {
            const Seg& u = itE.u;
            pp = s_pp[itE.ord & 3];
            col_ok = u.c0 + c < W;
            ys = y + (size_t)u.plane * H * W + (u.c0 + c);
            e_r0 = u.r0;
            e_len = u.seg_len;
        };
        auto epilogue = [&]() {  // block itE.b of the segment: outputs o = R b - 24 + 12 gq + i, i in [0, 12)
            mbar_wait_warp(&ctl.p2_done, epar);
            fence_after_sync();
            BTC_STAMP(ep_ts);
            float bl[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; i += 4) tmem_ld4_nowait(tmem + ((uint32_t)(32 * q) << 16) + L::OUT + (uint32_t)(L::RPW * gq + i), bl + i);
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive_warp(&ctl.out_empty);  // the accumulator is in registers: pass 2 of the next block may overwrite it
            const unsigned char* ctr = (gq == 0 ? side + (size_t)eside * L::SIDE_BYTES : tiles + (size_t)ebuf * L::T_BUF) + ctr_off;
            epar ^= 1u;
            ebuf = ebuf == kNBuf - 1 ? 0 : ebuf + 1;
            eside = (eside + 1) & (kSide - 1);
            const int o0 = R * itE.b - 2 * kRadius + L::RPW * gq;           // first output row of this warp
            const int i_lo = max(0, -o0), i_hi = min(L::RPW, e_len - o0);
            if (!col_ok || i_lo >= i_hi) return;
            // byte pointer advanced row by row (one 64-bit add per output; rows with i < i_lo are never dereferenced)
            char* yp = reinterpret_cast<char*>(ys) + (ptrdiff_t)(e_r0 + o0) * (ptrdiff_t)ystep_b;
            auto one = [&](int i) {
                float x4 = *reinterpret_cast<const float*>(ctr + 16 * i);
                if (X3 && gq != 0) x4 += *reinterpret_cast<const float*>(ctr + 16 * i + L::T_BYTES);
                Elem<TY>::store1_streaming(reinterpret_cast<TY*>(yp), DBG ? bl[i] : fmaf(x4 - bl[i], pp, x4));
                yp += ystep_b;
            };
            if (i_lo == 0 && i_hi == L::RPW) {  // the common case: every output row of the warp lies inside the segment
#pragma unroll
                for (int i = 0; i < L::RPW; ++i) one(i);
            } else {
#pragma unroll
                for (int i = 0; i < L::RPW; ++i) {
                    if (i >= i_lo && i < i_hi) one(i);
                    else yp += ystep_b;
                }
            }
        };

        // ---- the stream -------------------------------------------------------------------------------------------------------
        bool moreS = true;
        for (uint32_t g = 0; g < 2 && moreS; ++g) {  // prime: blocks 0 and 1
            moreS = itS.next(sc, H, entered);
            if (!moreS) break;
            if (entered) stage_enter();
            stage_load();
            stage_store();
        }
        uint32_t G = 0;
        while (itM.next(sc, H, entered)) {
            const int ts = (warp == 0 && G < 24) ? 16 + 8 * (int)G : -100;
            BTC_STAMP(ts);
            if (G > 0) {  // epilogue of block G - 1
                itE.next(sc, H, entered);
                if (entered) epilogue_enter();
                ep_ts = ts + 3;
                epilogue();
            }
            BTC_STAMP(ts + 4);
            compute_sync();  // every warp is done with block G-1's tile buffer and side slots before block G+2 is staged into them
            BTC_STAMP(ts + 5);
            if (moreS) moreS = itS.next(sc, H, entered);
            if (moreS) {
                if (entered) stage_enter();
                stage_load();
            }
            mbar_wait_warp(&ctl.p1_done[G % kNBuf], (G / kNBuf) & 1u);  // pass 1 of block G complete: ring columns valid
            BTC_STAMP(ts + 1);
            if (X3) {
                fence_after_sync();
                split_ring<L>(tmem, itM.b, warp);
                fence_before_sync();
                mbar_arrive_warp(&ctl.split_done);
            }
            BTC_STAMP(ts + 2);
            if (moreS) stage_store();
            BTC_STAMP(ts + 6);
            ++G;
        }
        if (G > 0) {
            itE.next(sc, H, entered);
            if (entered) epilogue_enter();
            ep_ts = -100;
            epilogue();
        }
    }

    teardown<L>(tmem);
}

}  // namespace btc
}  // namespace dd
