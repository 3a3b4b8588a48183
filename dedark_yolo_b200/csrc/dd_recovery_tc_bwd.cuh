// dd_recovery_tc_bwd.cuh -- a14: backward of the fused filter chain with the adjoint blur B^T g on the tensor cores
// (dd_blur_tc.cuh).  Same closed form, same partial-sum workspace and the same finalize kernel as dd_recovery_bwd.cu.
//
//     stage(b)     g (cotangent) -> mirror-extended g^ (reflected halo rows / columns, image-border row and column doubled),
//                  split, into tile buffer b % 3
//     pass 1, [split], pass 2   on the tensor cores: B^T g = blur of g^, twice the true value on the border row / column
//     epilogue(b)  thread <-> column: x0 (prefetched from global memory before the accumulator is awaited), chain recompute,
//                  the five parameter sums in registers, the row sums S = sum_w g4 x3 by warp shuffle (one partial per
//                  32-column quarter of a strip), optional dL/dx
#pragma once
#include "dd_blur_tc.cuh"

namespace dd {
namespace btc {

constexpr int kSpartPerStrip = 4;  // row-sum partials per (row, strip): one per 32-column quarter

// TX / TG: element types of x (and dx) and of the cotangent g (float, or __nv_bfloat16 in the bf16 I/O mode)
template <int R, bool X3, bool HAS_ICA, bool FAST, typename TX = float, typename TG = float>
__global__ void __launch_bounds__(kThreadsTC, 1)
recovery_bwd_tc_kernel(const TX* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                       const float* __restrict__ feat, const TG* __restrict__ g, float* __restrict__ part,
                       float* __restrict__ Spart, TX* __restrict__ dx, int B, int H, int W) {
    using L = Lay<R, X3>;
    using SG = StageGeo<L>;
    static_assert(R == 48, "the stage / epilogue work split assumes 48-row blocks (12 row groups x 4 rows, 12 outputs per warp)");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;
    unsigned char* side = tiles + (size_t)kNBuf * L::T_BUF;
    unsigned char* g_hi = side + (size_t)kSide * L::SIDE_BYTES;
    unsigned char* g_lo = g_hi + L::G_BYTES;
    __shared__ float MSm[kMaxU];  // per staged row of the segment in the epilogue: m = (1 - c) + c q   (0 outside the image)
    __shared__ float MSq[kMaxU];  // q - 1
    __shared__ ImgParams sp;
    __shared__ float s_red[kCW * kBwdSums];
    __shared__ Ctl ctl;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = setup<L>(ctl, g_hi, g_lo);   // reads no global memory: overlaps the tail of the previous kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const Sched sc = make_sched(B, H, W, kSchedCtasTC);
    const int Lc = W - 1;
    const int nsp = sc.strips * kSpartPerStrip;

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

        // ---- stage side: the cotangent needs no parameters, only geometry ---------------------------------------------------------
        SG sg;
        const TG* gp = g;
        int rowbase = 0, nU = 0;
        float fc[4];   // column factors of g^: the image-border columns are doubled in place; mirrored copies (halo chunks) are not
        float4 in[4];  // the block in flight between stage_load and stage_store
        auto stage_enter = [&]() {
            const Seg& u = itS.u;
            gp = g + (size_t)u.plane * H * W;
            rowbase = u.r0 - kRadius;
            nU = u.nU;
            sg.init(tid, u.c0, W);
#pragma unroll
            for (int i = 0; i < 4; ++i) fc[i] = (sg.gc + i == 0 || sg.gc + i == Lc) ? 2.f : 1.f;
        };
        auto stage_load = [&]() {
            if (sg.kind == SG::KIND_INSIDE || sg.kind == SG::KIND_MIRRORED) {
                const int v0 = R * itS.b + sg.rg;
#pragma unroll
                for (int k = 0; k < 4; ++k) in[k] = sg.load4(gp, (unsigned)reflect(rowbase + min(v0 + SG::NG * k, nU - 1), H) * (unsigned)W);
                if (sg.kind == SG::KIND_INSIDE && itS.b + 1 < itS.nB) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        prefetch_l2(gp + (unsigned)reflect(rowbase + min(v0 + R + SG::NG * k, nU - 1), H) * (unsigned)W + sg.gc);
                }
            }
        };
        auto stage_store = [&](uint32_t gg) {
            unsigned char* trow = tiles + (size_t)(gg % kNBuf) * L::T_BUF + sg.toff;
            const int v0 = R * itS.b + sg.rg;
            if (sg.kind == SG::KIND_INSIDE || sg.kind == SG::KIND_MIRRORED) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int v = v0 + SG::NG * k;
                    const int rho = rowbase + v;  // image row of this staged row (outside the image: mirrored, factor 1)
                    const float fr = v >= nU ? 0.f : (rho == 0 || rho == H - 1) ? 2.f : 1.f;  // rows past the segment: zeros
                    const float o[4] = {in[k].x * fr * fc[0], in[k].y * fr * fc[1], in[k].z * fr * fc[2], in[k].w * fr * fc[3]};
                    tile_store4<L>(trow + 16 * SG::NG * k, o);
                    if (k == 3 && sg.soff >= 0)
                        *reinterpret_cast<float4*>(side + (size_t)(gg % kSide) * L::SIDE_BYTES + sg.soff) = make_float4(o[0], o[1], o[2], o[3]);
                }
            } else if (sg.kind == SG::KIND_ZERO) {  // halo chunks beyond the reflect range
                const float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 4; ++k) tile_store4<L>(trow + 16 * SG::NG * k, z);
            }
            fence_proxy_async();
            mbar_arrive_warp(&ctl.tile_full[gg % kNBuf]);
        };

        // ---- epilogue side --------------------------------------------------------------------------------------------------------
        BwdAcc acc = {0.f, 0.f, 0.f, 0.f, 0.f};
        int cur_ps = -1;
        auto flush = [&]() {  // per (CTA, plane-strip) partial sums -> slot (cta + ps); fixed order: lanes, then warps by index
            float* out = part + (size_t)(blockIdx.x + cur_ps) * kBwdSums;
            const float v[kBwdSums] = {warp_sum(acc.p), warp_sum(acc.c), warp_sum(acc.g), warp_sum(acc.s), warp_sum(acc.w)};
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kBwdSums; ++j) s_red[warp * kBwdSums + j] = v[j];
            }
            compute_sync();
            if (tid < kBwdSums) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kCW; ++w) s += s_red[w * kBwdSums + tid];
                out[tid] = s;
            }
            compute_sync();
            acc.p = acc.c = acc.g = acc.s = acc.w = 0.f;
        };
        const int q = warp & 3, gq = warp >> 2, c = 32 * q + lane;   // lane quarter, output-row group (12 rows), strip column
        const int ctr_off = gq == 0 ? (c >> 2) * L::S_LBO + (c & 3) * 4 : tile_off<L>(L::RPW * (gq - 1), c + kRadius);
        ChainK ck;
        float pp = 0.f, hc = 1.f;
        bool col_ok = false;
        const TX* xcol = x;
        const float* icol = IcA;
        TX* dxcol = nullptr;
        float* spcol = Spart;
        int e_r0 = 0, e_len = 0;
        int ep_ts = -100;  // time-stamp slot of the epilogue in flight (DD_BTC_TIMING only)
        auto epilogue_enter = [&]() {  // first block of a new segment in the epilogue: its image's regressors, per-row contrast scalars
            const Seg& u = itE.u;
            if (cur_ps >= 0 && u.ps != cur_ps) flush();
            cur_ps = u.ps;
            compute_sync();  // every warp has finished the epilogues of the previous segment (MSm, MSq, sp are about to be rewritten)
            if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
            compute_sync();
            ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
            const float pc = sp.c;
            pp = sp.p;
            const TX* xp = x + (size_t)u.plane * H * W;
            const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;
            for (int v = tid; v < u.nU; v += kCT) {  // rows outside the image: 0
                const int row = u.r0 - kRadius + v;
                float m = 0.f, q1 = 0.f;
                if (row >= 0 && row < H) {
                    float x3[3];
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc)
                        x3[cc] = chain_x3<HAS_ICA, FAST>(ck, Elem<TX>::load1(xp + (size_t)row * W + cc), HAS_ICA ? __ldg(ip + (size_t)row * W + cc) : kDefaultIcA);
                    const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
                    q1 = rl.q - 1.f;
                    m = (1.f - pc) + pc * rl.q;
                }
                MSm[v] = m;
                MSq[v] = q1;
            }
            compute_sync();
            const int gcol = u.c0 + c;
            col_ok = gcol < W;
            hc = (gcol == 0 || gcol == Lc) ? 0.5f : 1.f;  // un-double the border column of g^ / halve the border output
            xcol = xp + gcol;
            icol = HAS_ICA ? ip + gcol : nullptr;
            dxcol = dx ? dx + (size_t)u.plane * H * W + gcol : nullptr;
            spcol = Spart + (size_t)u.plane * H * nsp + u.strip * kSpartPerStrip + q;
            e_r0 = u.r0;
            e_len = u.seg_len;
        };
        auto epilogue = [&](uint32_t gg) {  // block itE.b of the segment: outputs o = R b - 24 + 12 gq + i, i in [0, 12)
            const int o0 = R * itE.b - 2 * kRadius + L::RPW * gq;
            const int i_lo = max(0, -o0), i_hi = min(L::RPW, e_len - o0);
            const int row0 = e_r0 + o0;            // image row of output i = 0
            const int eoff = row0 * W;             // 32-bit element offset of that row
            // x0 (and IcA) of this thread's outputs: requested before the accumulator is awaited
            float x0v[L::RPW], icv[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; ++i) {
                x0v[i] = 0.f;
                icv[i] = kDefaultIcA;
                if (i >= i_lo && i < i_hi && col_ok) {
                    x0v[i] = Elem<TX>::load1(xcol + (eoff + i * W));
                    if (HAS_ICA) icv[i] = __ldg(icol + (eoff + i * W));
                }
            }
            mbar_wait_warp(&ctl.p2_done, gg & 1u);
            fence_after_sync();
            BTC_STAMP(ep_ts);
            float bt[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; i += 4) tmem_ld4_nowait(tmem + ((uint32_t)(32 * q) << 16) + L::OUT + (uint32_t)(L::RPW * gq + i), bt + i);
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive_warp(&ctl.out_empty);
            if (i_lo >= i_hi) return;
            const unsigned char* ctr = (gq == 0 ? side + (size_t)((gg + kSide - 1) % kSide) * L::SIDE_BYTES : tiles + (size_t)(gg % kNBuf) * L::T_BUF) + ctr_off;
            const float* msm = MSm + o0 + kRadius;
            const float* msq = MSq + o0 + kRadius;
#pragma unroll
            for (int i = 0; i < L::RPW; ++i) {
                if (i >= i_lo && i < i_hi) {  // warp-uniform
                    const int jr = row0 + i;  // image row
                    float gext = *reinterpret_cast<const float*>(ctr + 16 * i);
                    if (X3 && gq != 0) gext += *reinterpret_cast<const float*>(ctr + 16 * i + L::T_BYTES);
                    const float fac = (jr == 0 || jr == H - 1) ? 0.5f * hc : hc;
                    float srow = 0.f;
                    if (col_ok) {
                        const float d = px_bwd<HAS_ICA, FAST>(x0v[i], icv[i], gext * fac, bt[i] * fac, msm[i], msq[i], ck, pp, acc, srow);
                        if (dx) Elem<TX>::store1(dxcol + (eoff + i * W), d);
                    }
                    srow = warp_sum(srow);
                    if (lane == 0) spcol[(size_t)jr * nsp] = srow;
                }
            }
        };

        // ---- the stream -----------------------------------------------------------------------------------------------------------
        bool moreS = true;
        for (uint32_t gg = 0; gg < 2 && moreS; ++gg) {  // prime: blocks 0 and 1
            moreS = itS.next(sc, H, entered);
            if (!moreS) break;
            if (entered) stage_enter();
            stage_load();
            stage_store(gg);
        }
        uint32_t G = 0;
        while (itM.next(sc, H, entered)) {
            const int ts = (warp == 0 && G < 24) ? 16 + 8 * (int)G : -100;
            BTC_STAMP(ts);
            if (G > 0) {  // epilogue of block G - 1
                itE.next(sc, H, entered);
                if (entered) epilogue_enter();
                ep_ts = ts + 3;
                epilogue(G - 1);
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
            if (moreS) stage_store(G + 2);
            BTC_STAMP(ts + 6);
            ++G;
        }
        if (G > 0) {
            itE.next(sc, H, entered);
            if (entered) epilogue_enter();
            ep_ts = -100;
            epilogue(G - 1);
        }
        if (cur_ps >= 0) flush();
    }
    teardown<L>(tmem);
}

}  // namespace btc
}  // namespace dd
