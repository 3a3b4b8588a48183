// dd_conv_tiled.cuh -- CUDA-core kernels of the predictor's FIRST layer (3 -> 16 channels, 3x3 / stride 2 / pad 1, K = 27:
// too thin for the tensor cores, nn/modules/common.py:9-23): forward (+bias +LeakyReLU) and weight/bias gradient, fp32.
// (conv2..conv5 live in dd_conv_tc.cuh.)
//
// Both kernels are persistent (CTA c walks the tiles c, c + grid, ...) and their tiles arrive by TMA: one
// cp.async.bulk.tensor box per operand and tile, issued by one thread into one of two shared-memory buffers and counted
// on an mbarrier.  The box of the input starts at row 2*oh0 - 1 and column 2*ow0 - 4 (the innermost box origin must be
// 16-byte aligned -- an origin of -1 raises "illegal instruction", profiles/microbench/tma_probe.cu -- so the box carries
// three spare columns), and the convolution's zero padding is the copy engine's out-of-bounds fill: no thread spends an
// instruction on staging (the previous versions spent a third of theirs on 4-byte cp.async address arithmetic).
// Window access: lane = output column; taps kw = 1, 2 come from one aligned 64-bit shared load, kw = 0 from the
// previous lane (shuffle).
#pragma once
#include "dd_async.cuh"
#include "dd_common.cuh"

namespace dd {

constexpr int kC1In = 3, kC1Out = 16, kC1HIn = 256, kC1HOut = 128;
constexpr int kC1TW = 32;             // output columns per tile = lanes
constexpr int kC1BoxW = 72;           // input columns 2*ow0 - 4 .. 2*ow0 + 67 (3 spare on the left for alignment, 65 live, 4 pad)
__host__ __device__ constexpr int c1_tile_bytes(int th) { return ((kC1In * (2 * th + 1) * kC1BoxW * 4) + 127) / 128 * 128; }

// the three taps of one window row for output column `lane`
__device__ __forceinline__ void c1_window_row(const float* __restrict__ row, int lane, float& a, float& b, float& c) {
    const float2 bc = *reinterpret_cast<const float2*>(row + 2 * lane + 4);  // image columns 2*ow, 2*ow + 1
    b = bc.x;
    c = bc.y;
    a = __shfl_up_sync(0xffffffffu, bc.y, 1);                               // image column 2*ow - 1
    if (lane == 0) a = row[3];
}

// -------------------------------------------------------------------------------------------------------------
// forward: tile = 16 x 32 outputs x all 16 channels; thread = 4 rows x 1 column x 8 channels (256 threads)
// -------------------------------------------------------------------------------------------------------------
constexpr int kC1FwdTH = 16, kC1FwdPY = 4;
constexpr size_t conv1_fwd_smem() { return (size_t)2 * c1_tile_bytes(kC1FwdTH) + (size_t)kC1In * 9 * kC1Out * 4; }

__device__ __forceinline__ void conv1_fwd_body(const int cta, const int nctas, const CUtensorMap* rmap, const float* __restrict__ w,
                                               const float* __restrict__ bias, float* __restrict__ out, int ntiles,
                                               unsigned char* smem, uint64_t* full) {
    constexpr int TH = kC1FwdTH, PY = kC1FwdPY, TX = kC1HOut / kC1TW, TY = kC1HOut / TH, ROWS = 2 * TH + 1;
    constexpr int TILE_B = c1_tile_bytes(TH);
    float* s_w = reinterpret_cast<float*>(smem + 2 * TILE_B);  // [ci*9 + k][co]
    const int tid = threadIdx.x, lane = tid & 31, rg = (tid >> 5) & 3, cog = tid >> 7;
    for (int i = tid; i < kC1In * 9 * kC1Out; i += 256) s_w[i] = __ldg(w + (i % kC1Out) * (kC1In * 9) + i / kC1Out);
    auto issue = [&](int it) {
        const int tl = cta + it * nctas, tile = tl % (TX * TY), b = tl / (TX * TY);
        const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * kC1TW;
        mbar_arrive_expect_tx(&full[it & 1], kC1In * ROWS * kC1BoxW * 4);
        tma_load_3d(smem + (it & 1) * TILE_B, rmap, 2 * ow0 - 4, 2 * oh0 - 1, b * kC1In, &full[it & 1]);
    };
    const int my = cta < ntiles ? (ntiles - 1 - cta) / nctas + 1 : 0;
    if (tid == 0 && my > 0) issue(0);
    __syncthreads();  // weights
    for (int it = 0; it < my; ++it) {
        if (tid == 0 && it + 1 < my) issue(it + 1);  // its buffer was released by the barrier that ended iteration it - 1
        mbar_wait(&full[it & 1], (uint32_t)(it >> 1) & 1u);
        const float* s_in = reinterpret_cast<const float*>(smem + (it & 1) * TILE_B);
        float acc[PY][8];
#pragma unroll
        for (int j = 0; j < PY; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
#pragma unroll
        for (int ci = 0; ci < kC1In; ++ci) {
            const float* sp = s_in + (ci * ROWS + 2 * PY * rg) * kC1BoxW;
            float v[2 * PY + 1][3];
#pragma unroll
            for (int lr = 0; lr < 2 * PY + 1; ++lr) c1_window_row(sp + lr * kC1BoxW, lane, v[lr][0], v[lr][1], v[lr][2]);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4 wa = *reinterpret_cast<const float4*>(s_w + (ci * 9 + kh * 3 + kw) * kC1Out + 8 * cog);
                    const float4 wb = *reinterpret_cast<const float4*>(s_w + (ci * 9 + kh * 3 + kw) * kC1Out + 8 * cog + 4);
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
        const int tl = cta + it * nctas, tile = tl % (TX * TY), b = tl / (TX * TY);
        const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * kC1TW;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int co = 8 * cog + c;
            const float bv = __ldg(bias + co);
            float* o = out + (((size_t)b * kC1Out + co) * kC1HOut + oh0 + PY * rg) * kC1HOut + ow0 + lane;
#pragma unroll
            for (int j = 0; j < PY; ++j) o[(size_t)j * kC1HOut] = leaky(acc[j][c] + bv);
        }
        __syncthreads();  // everyone is done with this buffer before the next-but-one tile lands in it
    }
}

// -------------------------------------------------------------------------------------------------------------
// weight + bias gradient: tile = 8 x 32 output pixels; every thread keeps 4 x 27 (+4 bias) sums in registers across all of
// its CTA's tiles; ONE slice [16*27 + 16] per CTA at the end (summed in index order by the deferred reduction).
//   warp = (group of 4 output channels, half of the tile's rows), lane = output column.
// -------------------------------------------------------------------------------------------------------------
constexpr int kC1WgTH = 8;
constexpr int kC1DTileB = kC1Out * kC1WgTH * kC1TW * 4;  // 16 KB cotangent tile [co][row][32]
constexpr size_t conv1_wgrad_smem() { return (size_t)2 * (c1_tile_bytes(kC1WgTH) + kC1DTileB); }

__global__ void __launch_bounds__(256, 1)
conv1_wgrad_kernel(const __grid_constant__ CUtensorMap rmap, const __grid_constant__ CUtensorMap dmap, float* __restrict__ partial,
                   int ntiles) {
    constexpr int TH = kC1WgTH, TX = kC1HOut / kC1TW, TY = kC1HOut / TH, ROWS = 2 * TH + 1, NW = kC1Out * kC1In * 9;
    constexpr int IN_B = c1_tile_bytes(TH), BUF_B = IN_B + kC1DTileB;
    extern __shared__ __align__(128) unsigned char smem_c1[];
    __shared__ __align__(8) uint64_t full[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, cog = wid & 3, half = wid >> 2;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    // part 1 of a tile: the window of the resized batch r (an input of the forward pass); part 2: the cotangent tile the predecessor
    // (conv_tc_bwd<16, 32, 128>, which waits before it releases this kernel) has just written.  Both land on the same barrier.
    auto issue = [&](int it, int part) {
        const int tl = blockIdx.x + it * gridDim.x, tile = tl % (TX * TY), b = tl / (TX * TY);
        const int oh0 = (tile / TX) * TH, ow0 = (tile % TX) * kC1TW;
        unsigned char* buf = smem_c1 + (it & 1) * BUF_B;
        if (part & 1) {
            mbar_arrive_expect_tx(&full[it & 1], kC1In * ROWS * kC1BoxW * 4 + kC1DTileB);
            tma_load_3d(buf, &rmap, 2 * ow0 - 4, 2 * oh0 - 1, b * kC1In, &full[it & 1]);
        }
        if (part & 2) tma_load_3d(buf + IN_B, &dmap, ow0, oh0, b * kC1Out, &full[it & 1]);
    };
    const int my = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid == 0 && my > 0) issue(0, 1);   // r of the first tile: ahead of the grid dependency
    pdl_wait();
    pdl_launch();   // wait, then launch: most CTAs of the slice reduction behind this kernel do not wait for it (dd_conv_tc.cuh)
    if (tid == 0 && my > 0) issue(0, 2);

    float acc[4][27], accb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        accb[t] = 0.f;
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[t][k] = 0.f;
    }
    for (int it = 0; it < my; ++it) {
        if (tid == 0 && it + 1 < my) issue(it + 1, 3);
        mbar_wait(&full[it & 1], (uint32_t)(it >> 1) & 1u);
        const float* s_in = reinterpret_cast<const float*>(smem_c1 + (it & 1) * BUF_B);
        const float* s_d = reinterpret_cast<const float*>(smem_c1 + (it & 1) * BUF_B + IN_B);
#pragma unroll 4
        for (int j = 0; j < TH / 2; ++j) {
            const int oh = half * (TH / 2) + j;
            float v[27];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
                    c1_window_row(s_in + (ci * ROWS + 2 * oh + kh) * kC1BoxW, lane, v[ci * 9 + kh * 3], v[ci * 9 + kh * 3 + 1],
                                  v[ci * 9 + kh * 3 + 2]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float d = s_d[((cog * 4 + t) * TH + oh) * kC1TW + lane];
                accb[t] += d;
#pragma unroll
                for (int k = 0; k < 27; ++k) acc[t][k] = fmaf(d, v[k], acc[t][k]);
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled
    }
    // lanes -> one value per warp (fixed order), the two row halves through shared memory
    float* s_red = reinterpret_cast<float*>(smem_c1);  // [4 channel groups][112]
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
        float* o = partial + (size_t)blockIdx.x * (NW + kC1Out);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int co = cog * 4 + t;
#pragma unroll
            for (int k = 0; k < 27; ++k) o[co * 27 + k] = acc[t][k] + s_red[cog * 112 + t * 28 + k];  // (co*3 + ci)*9 + kk, k = ci*9 + kk
            o[NW + co] = accb[t] + s_red[cog * 112 + t * 28 + 27];
        }
    }
}

}  // namespace dd
