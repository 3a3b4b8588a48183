// dd_conv_tiled.cuh -- shared-memory-tiled 3x3 / stride 2 / pad 1 convolution kernels of the predictor
// (nn/modules/common.py:9-23): forward (+bias +LeakyReLU), data gradient, weight/bias gradient.  fp32, CUDA cores.
//
// Common idea: a CTA owns one spatial tile of one image; its operands are staged cooperatively in shared memory with
// all global loads in flight at once (the first version walked ~100 dependent L2 round trips per thread), and every
// thread keeps a register tile so that one shared-memory operand feeds 4..8 FMAs.  Input tiles are stored with the
// image columns split by parity ([even | odd] halves per row): a stride-2 convolution then reads consecutive
// addresses across the lanes of a warp (no bank conflicts).
#pragma once
#include "dd_common.cuh"

namespace dd {

// staged input tile: rows 2*oh0-1 .. 2*(oh0+TH-1)+1, columns 2*ow0-1 .. 2*(ow0+TW-1)+1 (zero outside the image).
// local column lc (0 .. 2TW) has image column 2*ow0 - 1 + lc: even lc -> odd image column, slot TW + lc/2;
// odd lc -> even image column, slot (lc-1)/2.  So for output column ow0 + c:  tap kw=0 -> odd slot TW + c,
// kw=1 -> even slot c, kw=2 -> odd slot TW + c + 1.
template <int TH, int TW>
struct InTile {
    static constexpr int ROWS = 2 * TH + 1;
    static constexpr int COLS = 2 * TW + 1;
    static constexpr int PITCH = 2 * TW + 2;
    static constexpr int PLANE = ROWS * PITCH + 1;  // odd plane stride: lanes that differ in channel hit different banks
};

// 4-byte async global->shared copy (LDGSTS), zero-fill when !valid.  All staging below is asynchronous: every
// chunk of every operand is requested up front (one commit group per chunk) and consumed as it lands.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_wait_dyn(int pending) {  // pending in 0..3 (compile-time after unrolling)
    switch (pending) {
        case 0: cp_wait<0>(); break;
        case 1: cp_wait<1>(); break;
        case 2: cp_wait<2>(); break;
        default: cp_wait<3>(); break;
    }
}

template <int TH, int TW, int NCH, int HIN>
__device__ __forceinline__ void stage_input_tile_async(float* __restrict__ s, const float* __restrict__ in_img, int c0,
                                                       int oh0, int ow0) {
    using T = InTile<TH, TW>;
    for (int idx = threadIdx.x; idx < NCH * T::ROWS * T::COLS; idx += blockDim.x) {
        const int lc = idx % T::COLS, t = idx / T::COLS, lr = t % T::ROWS, ci = t / T::ROWS;
        const int gy = 2 * oh0 - 1 + lr, gx = 2 * ow0 - 1 + lc;
        const bool ok = gy >= 0 && gy < HIN && gx >= 0 && gx < HIN;
        cp_async4(s + ci * T::PLANE + lr * T::PITCH + ((lc & 1) ? (lc >> 1) : TW + (lc >> 1)),
                  ok ? in_img + ((size_t)(c0 + ci) * HIN + gy) * HIN + gx : in_img, ok);
    }
}

template <int TH, int TW, int NCH, int HIN>
__device__ __forceinline__ void stage_input_tile(float* __restrict__ s, const float* __restrict__ in_img /* [C][HIN][HIN] of image b */,
                                                 int c0, int oh0, int ow0) {
    using T = InTile<TH, TW>;
    constexpr int N = NCH * T::ROWS * T::COLS, U = 8;
    // U loads are issued back to back before the first store (in-order issue would otherwise serialise every
    // load -> store pair on the L2 latency)
    for (int base = threadIdx.x; base < N; base += U * blockDim.x) {
        float v[U];
        int dst[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int idx = base + q * blockDim.x;
            const int lc = idx % T::COLS, t = idx / T::COLS, lr = t % T::ROWS, ci = t / T::ROWS;
            const int gy = 2 * oh0 - 1 + lr, gx = 2 * ow0 - 1 + lc;
            v[q] = 0.f;
            if (idx < N && gy >= 0 && gy < HIN && gx >= 0 && gx < HIN)
                v[q] = __ldg(in_img + ((size_t)(c0 + ci) * HIN + gy) * HIN + gx);
            dst[q] = ci * T::PLANE + lr * T::PITCH + ((lc & 1) ? (lc >> 1) : TW + (lc >> 1));
        }
#pragma unroll
        for (int q = 0; q < U; ++q)
            if (base + q * blockDim.x < N) s[dst[q]] = v[q];
    }
}

// -------------------------------------------------------------------------------------------------------------
// forward: thread = PY output rows x 1 output column x 8 output channels; CTA = TH x TW outputs x all COUT channels
// -------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int TH, int TW, int CICH, int PY>
__device__ __forceinline__ void conv_fwd_tiled_body(const int bid, const float* __restrict__ in, const float* __restrict__ w,
                                                    const float* __restrict__ bias, float* __restrict__ out) {
    using T = InTile<TH, TW>;
    constexpr int HO = HIN / 2, TX = HO / TW, TY = HO / TH, NCHUNK = CIN / CICH;
    constexpr int IN_F = (CICH * T::PLANE + 3) & ~3, W_F = CICH * 9 * COUT, CH_F = IN_F + W_F;  // floats per chunk
    extern __shared__ __align__(16) float smem_f[];
    const int tile = bid % (TX * TY), b = bid / (TX * TY);
    const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * TW;
    const int col = threadIdx.x % TW, rg = (threadIdx.x / TW) % (TH / PY), cog = threadIdx.x / (TW * (TH / PY));
    const float* in_img = in + (size_t)b * CIN * HIN * HIN;

#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        float* s_in = smem_f + ch * CH_F;
        float* s_w = s_in + IN_F;
        stage_input_tile_async<TH, TW, CICH, HIN>(s_in, in_img, ch * CICH, oh0, ow0);
        for (int idx = threadIdx.x; idx < W_F; idx += blockDim.x) {  // s_w[ci][k][co], co fastest
            const int co = idx % COUT, t = idx / COUT, k = t % 9, ci = t / 9;
            cp_async4(s_w + idx, w + ((size_t)co * CIN + ch * CICH + ci) * 9 + k, true);
        }
        cp_commit();
    }

    float acc[PY][8];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const float* s_in = smem_f + ch * CH_F;
        const float* s_w = s_in + IN_F;
        cp_wait_dyn(NCHUNK - 1 - ch);
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CICH; ++ci) {
            const float* sp = s_in + ci * T::PLANE + (2 * PY * rg) * T::PITCH;
            float v[2 * PY + 1][3];
#pragma unroll
            for (int lr = 0; lr < 2 * PY + 1; ++lr) {
                v[lr][0] = sp[lr * T::PITCH + TW + col];
                v[lr][1] = sp[lr * T::PITCH + col];
                v[lr][2] = sp[lr * T::PITCH + TW + col + 1];
            }
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 wa = *reinterpret_cast<const float4*>(s_w + (ci * 9 + kh * 3 + kw) * COUT + 8 * cog);
                    const float4 wb = *reinterpret_cast<const float4*>(s_w + (ci * 9 + kh * 3 + kw) * COUT + 8 * cog + 4);
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        const float a = v[2 * j + kh][kw];
                        acc[j][0] = fmaf(a, wa.x, acc[j][0]); acc[j][1] = fmaf(a, wa.y, acc[j][1]);
                        acc[j][2] = fmaf(a, wa.z, acc[j][2]); acc[j][3] = fmaf(a, wa.w, acc[j][3]);
                        acc[j][4] = fmaf(a, wb.x, acc[j][4]); acc[j][5] = fmaf(a, wb.y, acc[j][5]);
                        acc[j][6] = fmaf(a, wb.z, acc[j][6]); acc[j][7] = fmaf(a, wb.w, acc[j][7]);
                    }
                }
        }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int co = 8 * cog + c;
        const float bv = __ldg(bias + co);
        float* o = out + (((size_t)b * COUT + co) * HO + oh0 + PY * rg) * HO + ow0 + col;
#pragma unroll
        for (int j = 0; j < PY; ++j) o[(size_t)j * HO] = leaky(acc[j][c] + bv);
    }
}

// -------------------------------------------------------------------------------------------------------------
// data gradient.  A "quad" is the 2x2 block of input pixels (2a..2a+1, 2c..2c+1); the four parities of a stride-2 3x3
// conv touch exactly the 2x2 output neighbourhood (a..a+1, c..c+1):
//   (even,even): w11 d(a,c)                    (even,odd): w10 d(a,c+1) + w12 d(a,c)
//   (odd,even):  w01 d(a+1,c) + w21 d(a,c)     (odd,odd):  w00 d(a+1,c+1) + w02 d(a+1,c) + w20 d(a,c+1) + w22 d(a,c)
// thread = 2 vertically adjacent quads x 8 input channels; CTA = TQH x TQW quads x all CIN channels.  The result is
// multiplied by LeakyReLU'(act_in): the stored tensor is the gradient w.r.t. the previous layer's PRE-activation.
// -------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int TQH, int TQW, int COCH, int QY>
__device__ __forceinline__ void conv_dgrad_body(const int bid, const float* __restrict__ dpre, const float* __restrict__ w,
                                                const float* __restrict__ act_in, float* __restrict__ din) {
    constexpr int HO = HIN / 2, TX = HO / TQW, TY = HO / TQH, NCHUNK = COUT / COCH;
    constexpr int DP = TQW + 2, DPLANE = (TQH + 1) * DP;
    constexpr int D_F = (COCH * DPLANE + 3) & ~3, W_F = COCH * 9 * CIN, CH_F = D_F + W_F;  // floats per chunk
    extern __shared__ __align__(16) float smem_f[];
    const int tile = bid % (TX * TY), b = bid / (TX * TY);
    const int a0 = (tile / TX) * TQH, c0q = (tile % TX) * TQW;
    const int qc = threadIdx.x % TQW, qr = (threadIdx.x / TQW) % (TQH / QY), cig = threadIdx.x / (TQW * (TQH / QY));

#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        float* s_d = smem_f + ch * CH_F;
        float* s_w = s_d + D_F;
        for (int idx = threadIdx.x; idx < COCH * DPLANE; idx += blockDim.x) {
            const int lc = idx % DP, t = idx / DP, lr = t % (TQH + 1), co = t / (TQH + 1);
            const int oy = a0 + lr, ox = c0q + lc;
            const bool ok = lc <= TQW && oy < HO && ox < HO;
            cp_async4(s_d + idx, ok ? dpre + (((size_t)b * COUT + ch * COCH + co) * HO + oy) * HO + ox : dpre, ok);
        }
        for (int idx = threadIdx.x; idx < W_F; idx += blockDim.x) {  // [co][k][ci], ci fastest
            const int ci = idx % CIN, t = idx / CIN, k = t % 9, co = t / 9;
            cp_async4(s_w + idx, w + ((size_t)(ch * COCH + co) * CIN + ci) * 9 + k, true);
        }
        cp_commit();
    }

    float acc[QY][4][8];  // [quad][ee, eo, oe, oo][ci]
#pragma unroll
    for (int q = 0; q < QY; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][i][j] = 0.f;

#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const float* s_d = smem_f + ch * CH_F;
        const float* s_w = s_d + D_F;
        cp_wait_dyn(NCHUNK - 1 - ch);
        __syncthreads();
#pragma unroll 1
        for (int co = 0; co < COCH; ++co) {
            const float* dp = s_d + co * DPLANE + (QY * qr) * DP + qc;
            const float d00 = dp[0], d01 = dp[1], d10 = dp[DP], d11 = dp[DP + 1];
            const float d20 = QY > 1 ? dp[2 * DP] : 0.f, d21 = QY > 1 ? dp[2 * DP + 1] : 0.f;
            const float* wp = s_w + co * 9 * CIN + 8 * cig;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float4 wa = *reinterpret_cast<const float4*>(wp + k * CIN);
                const float4 wb = *reinterpret_cast<const float4*>(wp + k * CIN + 4);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                // tap k = kh*3+kw contributes to parity class (ph, pw) = (kh != 1, kw != 1) with output offset
                // (dy, dx) = (kh == 0, kw == 0)
                const int kh = k / 3, kw = k % 3;
                const int cls = (kh != 1 ? 2 : 0) + (kw != 1 ? 1 : 0);
                const bool dy = kh == 0, dx = kw == 0;
                const float dq0 = dy ? (dx ? d11 : d10) : (dx ? d01 : d00);  // quad 0 (rows a, a+1 of d)
                const float dq1 = dy ? (dx ? d21 : d20) : (dx ? d11 : d10);  // quad 1 (rows a+1, a+2 of d)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][cls][j] = fmaf(wv[j], dq0, acc[0][cls][j]);
                    if (QY > 1) acc[QY - 1][cls][j] = fmaf(wv[j], dq1, acc[QY - 1][cls][j]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < QY; ++q) {
        const int a = a0 + QY * qr + q, c = c0q + qc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ci = 8 * cig + j;
            const size_t base = (((size_t)b * CIN + ci) * HIN + 2 * a) * HIN + 2 * c;
            float2 top = make_float2(acc[q][0][j], acc[q][1][j]), bot = make_float2(acc[q][2][j], acc[q][3][j]);
            if (act_in) {
                const float2 at = *reinterpret_cast<const float2*>(act_in + base);
                const float2 ab = *reinterpret_cast<const float2*>(act_in + base + HIN);
                top.x = leaky_grad(at.x, top.x); top.y = leaky_grad(at.y, top.y);
                bot.x = leaky_grad(ab.x, bot.x); bot.y = leaky_grad(ab.y, bot.y);
            }
            *reinterpret_cast<float2*>(din + base) = top;
            *reinterpret_cast<float2*>(din + base + HIN) = bot;
        }
    }
}

// -------------------------------------------------------------------------------------------------------------
// weight + bias gradient, channel-parallel form (CIN >= 16): CTA = one TH x TW tile of output positions of one image;
// thread = (input channel ci, group of CO_T output channels) keeps CO_T x 9 sums over the tile's positions and writes
// them as one slice of the partial buffer [slice][COUT*CIN*9 + COUT]; wgrad_reduce_kernel adds the slices in index order.
// -------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int TH, int TW, int CO_T>
__device__ __forceinline__ void conv_wgrad_body(const int bid, const float* __restrict__ in, const float* __restrict__ dpre,
                                                float* __restrict__ partial) {
    using T = InTile<TH, TW>;
    constexpr int HO = HIN / 2, TX = HO / TW, TY = HO / TH, NW = COUT * CIN * 9;
    extern __shared__ __align__(16) float smem_w[];
    float* s_in = smem_w;                   // [CIN] planes, parity-split columns
    float* s_d = s_in + CIN * T::PLANE;     // [COUT][TH][TW]
    const int tile = bid % (TX * TY), b = bid / (TX * TY);
    const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * TW;
    stage_input_tile_async<TH, TW, CIN, HIN>(s_in, in + (size_t)b * CIN * HIN * HIN, 0, oh0, ow0);
    for (int idx = threadIdx.x; idx < COUT * TH * TW; idx += blockDim.x) {
        const int lc = idx % TW, t = idx / TW, lr = t % TH, co = t / TH;
        cp_async4(s_d + idx, dpre + (((size_t)b * COUT + co) * HO + oh0 + lr) * HO + ow0 + lc, true);
    }
    cp_commit();
    cp_wait<0>();
    __syncthreads();
    const int ci = threadIdx.x % CIN, cog = threadIdx.x / CIN;
    float acc[CO_T][9], accb[CO_T];
#pragma unroll
    for (int t = 0; t < CO_T; ++t) {
        accb[t] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[t][k] = 0.f;
    }
    const float* sp = s_in + ci * T::PLANE;
#pragma unroll 1
    for (int oh = 0; oh < TH; ++oh)
#pragma unroll 2
        for (int ow = 0; ow < TW; ++ow) {
            float v[9];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* rp = sp + (2 * oh + kh) * T::PITCH;
                v[kh * 3 + 0] = rp[TW + ow];
                v[kh * 3 + 1] = rp[ow];
                v[kh * 3 + 2] = rp[TW + ow + 1];
            }
#pragma unroll
            for (int t = 0; t < CO_T; ++t) {
                const float d = s_d[((cog * CO_T + t) * TH + oh) * TW + ow];
                accb[t] += d;
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[t][k] = fmaf(d, v[k], acc[t][k]);
            }
        }
    float* out = partial + (size_t)bid * (NW + COUT);
#pragma unroll
    for (int t = 0; t < CO_T; ++t) {
        const int co = cog * CO_T + t;
#pragma unroll
        for (int k = 0; k < 9; ++k) out[(co * CIN + ci) * 9 + k] = acc[t][k];
        if (ci == 0) out[NW + co] = accb[t];
    }
}

// position-parallel form for the first layer (CIN = 3): thread = (lane = slice of the tile's positions, warp = group of
// 2 output channels) keeps 2 x 27 sums, lanes are combined with warp shuffles in a fixed order.
template <int COUT, int HIN, int TH, int TW>
__global__ void __launch_bounds__(32 * (COUT / 2))
conv_wgrad_tiled_c3(const float* __restrict__ in, const float* __restrict__ dpre, float* __restrict__ partial) {
    pdl_begin();
    constexpr int CIN = 3;
    using T = InTile<TH, TW>;
    constexpr int HO = HIN / 2, TX = HO / TW, TY = HO / TH, NW = COUT * CIN * 9;
    __shared__ float s_in[CIN * T::PLANE];
    __shared__ float s_d[COUT * TH * TW];
    const int tile = blockIdx.x % (TX * TY), b = blockIdx.x / (TX * TY);
    const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * TW;
    stage_input_tile_async<TH, TW, CIN, HIN>(s_in, in + (size_t)b * CIN * HIN * HIN, 0, oh0, ow0);
    for (int idx = threadIdx.x; idx < COUT * TH * TW; idx += blockDim.x) {
        const int lc = idx % TW, t = idx / TW, lr = t % TH, co = t / TH;
        cp_async4(s_d + idx, dpre + (((size_t)b * COUT + co) * HO + oh0 + lr) * HO + ow0 + lc, true);
    }
    cp_commit();
    cp_wait<0>();
    __syncthreads();
    const int lane = threadIdx.x & 31, cog = threadIdx.x >> 5;
    float acc[2][27], accb[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[t][k] = 0.f;
#pragma unroll 1
    for (int p = lane; p < TH * TW; p += 32) {
        const int oh = p / TW, ow = p % TW;
        float v[27];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* rp = s_in + ci * T::PLANE + (2 * oh + kh) * T::PITCH;
                v[ci * 9 + kh * 3 + 0] = rp[TW + ow];
                v[ci * 9 + kh * 3 + 1] = rp[ow];
                v[ci * 9 + kh * 3 + 2] = rp[TW + ow + 1];
            }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const float d = s_d[((cog * 2 + t) * TH + oh) * TW + ow];
            accb[t] += d;
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[t][k] = fmaf(d, v[k], acc[t][k]);
        }
    }
    float* out = partial + (size_t)blockIdx.x * (NW + COUT);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int co = cog * 2 + t;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float s = warp_sum(acc[t][k]);
            if (lane == 0) out[co * 27 + k] = s;  // (co*3 + ci)*9 + kk with k = ci*9 + kk
        }
        const float sb = warp_sum(accb[t]);
        if (lane == 0) out[NW + co] = sb;
    }
}

// One launch per layer of the backward: the weight gradient (CTAs 0 .. n_wgrad-1) and the data gradient (the rest) of a
// layer are independent given dpre, so they share a grid (more CTAs per launch, half the dependent launches).
template <int CIN, int COUT, int HIN, int WTH, int WTW, int CO_T, int TQH, int TQW, int COCH, int QY>
__global__ void __launch_bounds__(256)
conv_bwd_layer(const float* __restrict__ in, const float* __restrict__ dpre, const float* __restrict__ w,
               const float* __restrict__ act_in, float* __restrict__ partial, float* __restrict__ din, int n_wgrad) {
    pdl_begin();
    static_assert(CIN * (COUT / CO_T) == 256 && TQW * (TQH / QY) * (CIN / 8) == 256, "both halves run with 256 threads");
    if ((int)blockIdx.x < n_wgrad)
        conv_wgrad_body<CIN, COUT, HIN, WTH, WTW, CO_T>(blockIdx.x, in, dpre, partial);
    else
        conv_dgrad_body<CIN, COUT, HIN, TQH, TQW, COCH, QY>(blockIdx.x - n_wgrad, dpre, w, act_in, din);
}

// all five layers' slice sums in one launch (deferred to the end of the backward)
struct ReduceJob {
    const float* partial;
    float* dw;
    float* db;
    int split, nw, nb, block0;  // block0: first CTA of this job
};
struct ReduceJobs {
    ReduceJob j[5];
};
__global__ void __launch_bounds__(1024) wgrad_reduce_all_kernel(const ReduceJobs jobs) {
    pdl_begin();
    __shared__ float s_part[32][33];
    int l = 4;
#pragma unroll
    for (int q = 3; q >= 0; --q)
        if ((int)blockIdx.x < jobs.j[q + 1].block0) l = q;
    const ReduceJob jb = jobs.j[l];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = jb.nw + jb.nb;
    const int i = ((int)blockIdx.x - jb.block0) * 32 + lane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (i < n) {
        int s = wid;
        for (; s + 96 < jb.split; s += 128) {
            a0 += __ldg(jb.partial + (size_t)s * n + i);
            a1 += __ldg(jb.partial + (size_t)(s + 32) * n + i);
            a2 += __ldg(jb.partial + (size_t)(s + 64) * n + i);
            a3 += __ldg(jb.partial + (size_t)(s + 96) * n + i);
        }
        for (; s < jb.split; s += 32) a0 += __ldg(jb.partial + (size_t)s * n + i);
    }
    s_part[wid][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (wid == 0 && i < n) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) r += s_part[k][lane];
        if (i < jb.nw) jb.dw[i] = r; else jb.db[i - jb.nw] = r;
    }
}

template <int CIN, int COUT, int TH, int TW, int CICH>
constexpr size_t conv_fwd_smem() {
    return (size_t)(CIN / CICH) * (((CICH * InTile<TH, TW>::PLANE + 3) & ~3) + CICH * 9 * COUT) * sizeof(float);
}
template <int CIN, int COUT, int TQH, int TQW, int COCH>
constexpr size_t conv_dgrad_smem() {
    return (size_t)(COUT / COCH) * (((COCH * (TQH + 1) * (TQW + 2) + 3) & ~3) + COCH * 9 * CIN) * sizeof(float);
}

template <int CIN, int COUT, int TH, int TW>
constexpr size_t conv_wgrad_smem() {
    return (size_t)(CIN * InTile<TH, TW>::PLANE + COUT * TH * TW) * sizeof(float);
}

}  // namespace dd
