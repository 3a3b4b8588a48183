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

template <int R, bool X3, bool HAS_ICA, bool FAST>
__global__ void __launch_bounds__(kThreadsTC, 1)
recovery_bwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ IcA,
                       const float* __restrict__ feat, const float* __restrict__ g, float* __restrict__ part,
                       float* __restrict__ Spart, float* __restrict__ dx, int B, int H, int W) {
    using L = Lay<R, X3>;
    pdl_begin();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;
    float* side = reinterpret_cast<float*>(tiles + (size_t)kNBuf * L::T_BUF);
    unsigned char* g_hi = reinterpret_cast<unsigned char*>(side) + (size_t)kSide * L::SIDE_BYTES;
    unsigned char* g_lo = g_hi + L::G_BYTES;
    __shared__ float MSm[kMaxU];  // per staged row: m = (1 - c) + c q   (0 outside the image)
    __shared__ float MSq[kMaxU];  // per staged row: q - 1
    __shared__ ImgParams sp;
    __shared__ float s_red[kCW * kBwdSums];
    __shared__ Ctl ctl;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = setup<L>(ctl, g_hi, g_lo);

    const Sched sc = make_sched(B, H, W, kSchedCtasTC);
    const long long blk_end = sched_begin(sc, blockIdx.x + 1);
    const int Lc = W - 1;
    const int nsp = sc.strips * kSpartPerStrip;
    uint32_t gb = 0;

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

    for (long long blk = sched_begin(sc, blockIdx.x); blk < blk_end;) {
        const Seg u = next_seg(blk, blk_end, sc, H);
        blk += seg_blocks(u);
        const int nB = (u.nU + R - 1) / R;

        if (warp == kCW) {
            mma_warp_segment<L>(ctl, tmem, tiles, smem_u32(g_hi), smem_u32(g_lo), nB, gb);
            continue;
        }

        // ================================ compute warps ================================
        if (cur_ps >= 0 && u.ps != cur_ps) flush();
        cur_ps = u.ps;
        compute_sync();  // the previous segment is fully drained
        if (tid < 32) regress_warp(feat + u.b * kFeat, sp);
        compute_sync();
        const ChainK ck = make_chain(sp, u.ch, A ? __ldg(A + u.b * 3 + u.ch) : kDefaultA);
        const float pc = sp.c, pp = sp.p;
        const float* xp = x + (size_t)u.plane * H * W;
        const float* gp = g + (size_t)u.plane * H * W;
        const float* ip = HAS_ICA ? IcA + (size_t)u.b * H * W : nullptr;

        for (int v = tid; v < u.nU; v += kCT) {  // per-row contrast scalars (rows outside the image: 0)
            const int row = u.r0 - kRadius + v;
            float m = 0.f, q1 = 0.f;
            if (row >= 0 && row < H) {
                float x3[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    x3[c] = chain_x3<HAS_ICA, FAST>(ck, __ldg(xp + (size_t)row * W + c), HAS_ICA ? __ldg(ip + (size_t)row * W + c) : kDefaultIcA);
                const RowLum rl = row_lum<false>(x3[0], x3[1], x3[2]);
                q1 = rl.q - 1.f;
                m = (1.f - pc) + pc * rl.q;
            }
            MSm[v] = m;
            MSq[v] = q1;
        }
        compute_sync();

        auto stage = [&](int b) {
            const uint32_t gg = gb + (uint32_t)b;
            unsigned char* tile = tiles + (size_t)(gg % kNBuf) * L::T_BUF;
            float* sd = side + (size_t)(gg % kSide) * (L::SIDE_BYTES / 4);
#pragma unroll 1
            for (int idx = tid; idx < R * kCH; idx += kCT) {
                const int r = idx / kCH, j = idx - r * kCH;
                const int v = R * b + r;
                const int gc = u.c0 - kRadius + 4 * j;
                const bool inside = gc >= 0 && gc < W, vrow = v < u.nU;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (inside && vrow) {
                    const int rho = u.r0 - kRadius + v;  // image row of this staged row (outside the image: mirrored)
                    const int row = reflect(rho, H);
                    const float fr = (rho == 0 || rho == H - 1) ? 2.f : 1.f;
                    const float4 in = __ldg(reinterpret_cast<const float4*>(gp + (size_t)row * W + gc));
                    const float raw[4] = {in.x * fr, in.y * fr, in.z * fr, in.w * fr};
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[i] = (gc + i == 0 || gc + i == Lc) ? 2.f * raw[i] : raw[i];
                    mirror_cols<L>(tile, r, gc, u.c0, Lc, raw);
                }
                if (inside || !vrow || !chunk_is_mirrored(gc, W)) tile_store4<L>(tile, r, j, o);
                if (r >= R - kRadius && j >= kRadius / 4 && j < kRadius / 4 + kStripW / 4)
                    *reinterpret_cast<float4*>(sd + (r - (R - kRadius)) * kStripW + 4 * j - kRadius) = make_float4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            mbar_arrive(&ctl.tile_full[gg % kNBuf]);
        };

        auto epilogue = [&](int bb) {  // outputs o = R * bb - 24 + n, n in [0, R)
            const uint32_t gg = gb + (uint32_t)bb;
            const int q = warp & 3, n_first = (warp >> 2) * L::RPW, c = 32 * q + lane;
            const int gc = u.c0 + c;
            const bool col_ok = gc < W;
            // x0 (and IcA) of this thread's outputs: requested before the accumulator is awaited
            float x0v[L::RPW], icv[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; ++i) {
                const int o = R * bb - 2 * kRadius + n_first + i;
                x0v[i] = 0.f;
                icv[i] = kDefaultIcA;
                if (o >= 0 && o < u.seg_len && col_ok) {
                    const size_t off = (size_t)(u.r0 + o) * W + gc;
                    x0v[i] = __ldg(xp + off);
                    if (HAS_ICA) icv[i] = __ldg(ip + off);
                }
            }
            mbar_wait(&ctl.p2_done, gg & 1u);
            fence_after_sync();
            float bt[L::RPW];
#pragma unroll
            for (int i = 0; i < L::RPW; i += 4) tmem_ld4_nowait(tmem + ((uint32_t)(32 * q) << 16) + L::OUT + (uint32_t)(n_first + i), bt + i);
            tmem_ld_wait();
            fence_before_sync();
            mbar_arrive(&ctl.out_empty);
            const unsigned char* tile = tiles + (size_t)(gg % kNBuf) * L::T_BUF;
            const float* sd = side + (size_t)((gg + kSide - 1) % kSide) * (L::SIDE_BYTES / 4);
            const float hc = (gc == 0 || gc == Lc) ? 0.5f : 1.f;  // un-double the border column (g^) / halve the border output
#pragma unroll
            for (int i = 0; i < L::RPW; ++i) {
                const int o = R * bb - 2 * kRadius + n_first + i;
                if (o < 0 || o >= u.seg_len) continue;  // warp-uniform
                const int jr = u.r0 + o;                 // image row
                const int rr = o + kRadius - R * bb;
                const float gext = rr < 0 ? sd[(rr + kRadius) * kStripW + c] : tile_load1<L>(tile, rr, c + kRadius);
                const float fac = ((jr == 0 || jr == H - 1) ? 0.5f : 1.f) * hc;
                float srow = 0.f;
                if (col_ok) {
                    const float d = px_bwd<HAS_ICA, FAST>(x0v[i], icv[i], gext * fac, bt[i] * fac, MSm[o + kRadius], MSq[o + kRadius], ck, pp, acc, srow);
                    if (dx) dx[(size_t)u.plane * H * W + (size_t)jr * W + gc] = d;
                }
                srow = warp_sum(srow);
                if (lane == 0) Spart[((size_t)u.plane * H + jr) * nsp + u.strip * kSpartPerStrip + q] = srow;
            }
        };

        stage(0);
        if (nB > 1) stage(1);
        for (int b = 0; b < nB; ++b) {
            const uint32_t gg = gb + (uint32_t)b;
            mbar_wait(&ctl.p1_done[gg % kNBuf], (gg / kNBuf) & 1u);
            if (X3) {
                fence_after_sync();
                split_ring<L>(tmem, b, warp);
                fence_before_sync();
                mbar_arrive(&ctl.split_done);
            }
            if (b > 0) epilogue(b - 1);
            compute_sync();
            if (b + 2 < nB) stage(b + 2);
        }
        epilogue(nB - 1);
        gb += (uint32_t)nB;
    }
    if (warp < kCW && cur_ps >= 0) flush();
    teardown<L>(tmem);
}

}  // namespace btc
}  // namespace dd
