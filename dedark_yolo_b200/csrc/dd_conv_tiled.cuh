// dd_conv_tiled.cuh -- shared-memory-tiled CUDA-core kernels of the predictor's FIRST layer (3 -> 16 channels, K = 27:
// too thin for the tensor cores, nn/modules/common.py:9-23): forward (+bias +LeakyReLU) and weight/bias gradient.  fp32.
// (conv2..conv5 live in dd_conv_tc.cuh.)
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


template <int CIN, int COUT, int TH, int TW, int CICH>
constexpr size_t conv_fwd_smem() {
    return (size_t)(CIN / CICH) * (((CICH * InTile<TH, TW>::PLANE + 3) & ~3) + CICH * 9 * COUT) * sizeof(float);
}

// -------------------------------------------------------------------------------------------------------------
// weight + bias gradient of the first layer (CIN = 3), persistent: CTA c walks the tiles c, c + grid, ... (a tile =
// TH x TW output pixels of one image), keeping its 4 x 27 (+4 bias) sums per thread in registers across tiles, and
// writes ONE slice [COUT*27 + COUT] at the end (summed in index order by the deferred reduction).
//   warp = (group of 4 output channels, half of the tile's rows), lane = output column: per pixel 27 window loads +
//   4 cotangent loads feed 112 FMAs.  Input and cotangent tiles are double-buffered with cp.async.
// -------------------------------------------------------------------------------------------------------------
template <int COUT, int TH, int TW>
constexpr size_t conv_wgrad_c3_smem() {
    return (size_t)2 * (((3 * InTile<TH, TW>::PLANE + 3) & ~3) + COUT * TH * TW) * sizeof(float);
}

template <int COUT, int HIN, int TH, int TW>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_c3_kernel(const float* __restrict__ in, const float* __restrict__ dpre, float* __restrict__ partial, int ntiles) {
    pdl_begin();
    constexpr int CIN = 3;
    using T = InTile<TH, TW>;
    constexpr int HO = HIN / 2, TX = HO / TW, TY = HO / TH, NW = COUT * CIN * 9;
    constexpr int IN_F = (CIN * T::PLANE + 3) & ~3, D_F = COUT * TH * TW, BUF = IN_F + D_F;
    static_assert(COUT == 16 && TW == 32 && TH == 8, "warp mapping: 4 channel groups x 2 row halves, lane = column");
    extern __shared__ __align__(16) float smem_w[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, cog = wid & 3, half = wid >> 2;

    auto stage = [&](int buf, int tl) {
        const int tile = tl % (TX * TY), b = tl / (TX * TY);
        const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * TW;
        float* s_in = smem_w + buf * BUF;
        float* s_d = s_in + IN_F;
        stage_input_tile_async<TH, TW, CIN, HIN>(s_in, in + (size_t)b * CIN * HIN * HIN, 0, oh0, ow0);
        for (int idx = threadIdx.x; idx < D_F / 4; idx += blockDim.x) {  // [co][row][TW] rows of 32 contiguous floats
            const int c4 = idx % (TW / 4), t = idx / (TW / 4), lr = t % TH, co = t / TH;
            const unsigned d = (unsigned)__cvta_generic_to_shared(s_d + 4 * idx);
            const float* src = dpre + (((size_t)b * COUT + co) * HO + oh0 + lr) * HO + ow0 + 4 * c4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
        }
        cp_commit();
    };

    float acc[4][27], accb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        accb[t] = 0.f;
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[t][k] = 0.f;
    }
    int buf = 0;
    if ((int)blockIdx.x < ntiles) stage(0, blockIdx.x);
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const bool more = tl + (int)gridDim.x < ntiles;
        if (more) stage(buf ^ 1, tl + gridDim.x);
        if (more) cp_wait<1>(); else cp_wait<0>();
        __syncthreads();
        const float* s_in = smem_w + buf * BUF;
        const float* s_d = s_in + IN_F;
#pragma unroll 1
        for (int j = 0; j < TH / 2; ++j) {
            const int oh = half * (TH / 2) + j;
            float v[27];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const float* rp = s_in + ci * T::PLANE + (2 * oh + kh) * T::PITCH;
                    v[ci * 9 + kh * 3 + 0] = rp[TW + lane];
                    v[ci * 9 + kh * 3 + 1] = rp[lane];
                    v[ci * 9 + kh * 3 + 2] = rp[TW + lane + 1];
                }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float d = s_d[((cog * 4 + t) * TH + oh) * TW + lane];
                accb[t] += d;
#pragma unroll
                for (int k = 0; k < 27; ++k) acc[t][k] = fmaf(d, v[k], acc[t][k]);
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
        buf ^= 1;
    }
    // lanes -> one value per warp (fixed order), the two row halves through shared memory
    float* s_red = smem_w;  // [4 channel groups][112]
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float sv = warp_sum(acc[t][k]);
            if (half == 1 && lane == 0) s_red[cog * 112 + t * 28 + k] = sv;
            acc[t][k] = sv;
        }
        const float sb = warp_sum(accb[t]);
        if (half == 1 && lane == 0) s_red[cog * 112 + t * 28 + 27] = sb;
        accb[t] = sb;
    }
    __syncthreads();
    if (half == 0 && lane == 0) {
        float* out = partial + (size_t)blockIdx.x * (NW + COUT);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int co = cog * 4 + t;
#pragma unroll
            for (int k = 0; k < 27; ++k) out[co * 27 + k] = acc[t][k] + s_red[cog * 112 + t * 28 + k];  // (co*3 + ci)*9 + kk, k = ci*9 + kk
            out[NW + co] = accb[t] + s_red[cog * 112 + t * 28 + 27];
        }
    }
}

}  // namespace dd
