// dd_predictor_tail.cuh -- the END of the parameter predictor's forward (conv4, conv5, fc1, fc2) as ONE launch.
//
// Reference: nn/modules/common.py:9-23 (ConvBlock: Conv2d k3 s2 p1 + LeakyReLU 0.1), :52-78 (ExtractParameters2: the last
// two conv blocks 32x32x32 -> 32x16x16 -> 32x8x8, flatten, fc1 2048->64 + LeakyReLU, fc2 64->15).
//
// EXPERIMENT, off by default (DEDARK_TAIL=fused selects it).  These layers are 47 MMAC for a batch of 16 -- microseconds of
// arithmetic -- but as three launches (two tcgen05 GEMMs with 32 and 8 tiles and the FC cluster kernel) each pays the fixed
// latency of its own pipeline (ncu, each kernel alone: 10 + 10 + 11 us, sm__cycles_active max ~12k cycles regardless of
// the tile count).  Here one CLUSTER of 8 CTAs owns one image:
//   phase A  conv4: CTA q computes output rows 2q, 2q+1 (32 pixels x 32 channels) from 5 input rows staged in shared memory
//            (thread = output channel [lane] x 4 neighbouring pixels [warp]; inputs are warp-wide broadcasts, weights are
//            conflict-free column reads of the transposed weight tile), writes a3 (the backward needs it)
//   phase B  conv5: CTA q computes output row q (8 pixels x 32 channels) from a3 rows 2q-1..2q+1 read back through L2
//            after a cluster barrier (thread = output channel x 1 pixel), writes a4 (= the flattened fc1 input)
//   phase C  fc1: warp w of CTA q computes h[8q + w] (its weight row is requested before phase B so that the loads fly
//            during the barrier and conv5); the eight h-octets meet in CTA 0 through distributed shared memory; fc2 there.
// All sums are serial fp32 FMA chains in a fixed order: bit-reproducible, and closer to the fp64 oracle than the 3xTF32
// GEMMs (act4 1.6e-6 vs 2.1e-6, feat 0.9e-7 vs 1.7e-7 rel-to-max).
// MEASURED (B = 16, bursts of back-to-back dd_predictor_fwd): 54.0 us with this tail, 51.3 us with the three launches.  In the
// stream the three kernels overlap each other's prologues through programmatic dependent launch (the sum of their isolated
// times is not what the step pays), while this kernel is one serial chain of five latency-bound phases (weight transposition
// 2 x 36 KB per CTA, staging, conv4, barrier, conv5, barrier, fc1, barrier, fc2) on 8 warps per SM: 29 us alone under ncu
// (issue slots 19 % busy).  Kept as a record and as an independent fp32 check of the tensor-core path (tests/test_gpu_tail.py).
#pragma once
#include <cooperative_groups.h>

#include "dd_common.cuh"
#include "dd_layout.cuh"

namespace dd {

constexpr int kTailCluster = 8;
constexpr int kTailWP = 33;          // pitch of the transposed weight tiles [ci*9 + k][co]: conflict-free fill and reads
constexpr int kTailInP = 36;         // pitch of a staged conv4 input row: column c at [c + 1], [0] = the left zero pad
constexpr int kTailIn5P = 20;        // the same for conv5 (16 columns)
constexpr int kTailWElems = 288 * kTailWP;
constexpr int kTailInElems = 32 * 5 * kTailInP;
constexpr int kTailIn5Elems = 32 * 3 * kTailIn5P;
constexpr size_t kTailSmem = (size_t)(2 * kTailWElems + kTailInElems + kTailIn5Elems) * sizeof(float);

inline bool tail_fused() {
    const char* e = getenv("DEDARK_TAIL");
    return e && e[0] == 'f';
}

// w [co][ci][3][3] (32 x 32) -> ws[(ci*9 + k) * 33 + co]: coalesced global reads (a warp reads 32 consecutive weights of one
// output channel), scattered down the transposed tile (stride 33: conflict-free), 12 loads in flight per thread.  (Gathering
// with lane = co and float4 loads -- fewer instructions, 32 sectors per request -- measured 7 us SLOWER.)
__device__ __forceinline__ void tail_load_weights(const float* __restrict__ w, float* __restrict__ ws, int tid) {
#pragma unroll 12
    for (int i = tid; i < 32 * 288; i += 256) {
        const int co = i / 288, r = i - co * 288;
        ws[r * kTailWP + co] = __ldg(w + i);
    }
}

__global__ void __launch_bounds__(256)
predictor_tail_kernel(const float* __restrict__ a2, const float* __restrict__ w4, const float* __restrict__ b4,
                      const float* __restrict__ w5, const float* __restrict__ b5, const float* __restrict__ w1,
                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                      float* __restrict__ a3, float* __restrict__ a4, float* __restrict__ h, float* __restrict__ feat) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float smem[];
    float* s_w4 = smem;
    float* s_w5 = s_w4 + kTailWElems;
    float* s_in = s_w5 + kTailWElems;
    float* s_in5 = s_in + kTailInElems;
    __shared__ float s_h[8];
    __shared__ float s_all[kFc1Out];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.x / kTailCluster, q = (int)cluster.block_rank();

    // the weights are parameters: not written by the kernel this one may overlap with, so their staging runs BEFORE the
    // grid dependency is resolved (dd_common.cuh pdl_begin: everything older than the predecessor has completed)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    tail_load_weights(w4, s_w4, tid);
    tail_load_weights(w5, s_w5, tid);
    for (int i = tid; i < 32 * 5; i += 256) s_in[i * kTailInP] = 0.f;        // left pads
    for (int i = tid; i < 32 * 3; i += 256) s_in5[i * kTailIn5P] = 0.f;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // ---- phase A: conv4 -----------------------------------------------------------------------------------------------
    {
        // input rows 4q-1 .. 4q+3 of all 32 channels (row -1 = zero pad)
        const float* src = a2 + (size_t)b * 32 * 32 * 32;
        for (int i = tid; i < 32 * 5 * 8; i += 256) {
            const int c4 = i & 7, t = i >> 3, lr = t % 5, ci = t / 5, row = 4 * q - 1 + lr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) v = __ldcg(reinterpret_cast<const float4*>(src + ((size_t)ci * 32 + row) * 32) + c4);  // producer data: never __ldg behind a wait
            float* d = s_in + (ci * 5 + lr) * kTailInP + 1 + 4 * c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        const int orow = wid >> 2, ow0 = (wid & 3) * 4;   // warp: 4 pixels of one output row; lane: output channel
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int ci = 0; ci < 32; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* ip = s_in + (ci * 5 + 2 * orow + kh) * kTailInP + 2 * ow0;
                const float4 v0 = *reinterpret_cast<const float4*>(ip), v1 = *reinterpret_cast<const float4*>(ip + 4);
                const float x[9] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, ip[8]};
                const float* wp = s_w4 + (ci * 9 + kh * 3) * kTailWP + lane;
                const float k0 = wp[0], k1 = wp[kTailWP], k2 = wp[2 * kTailWP];
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    acc[p] = fmaf(k2, x[2 * p + 2], fmaf(k1, x[2 * p + 1], fmaf(k0, x[2 * p], acc[p])));
            }
        }
        const float bias = __ldg(b4 + lane);
        float4 o = make_float4(leaky(acc[0] + bias), leaky(acc[1] + bias), leaky(acc[2] + bias), leaky(acc[3] + bias));
        *reinterpret_cast<float4*>(a3 + (((size_t)b * 32 + lane) * 16 + 2 * q + orow) * 16 + ow0) = o;
    }
    // fc1 weight row of this warp: requested now, consumed in phase C
    const int o1 = q * 8 + wid;
    float4 wrow[kFc1In / 128];
    {
        const float4* wa = reinterpret_cast<const float4*>(w1 + (size_t)o1 * kFc1In);
#pragma unroll
        for (int i = 0; i < kFc1In / 128; ++i) wrow[i] = __ldg(wa + lane + 32 * i);
    }
    cluster.sync();   // release/acquire at cluster scope: the peers' a3 / a4 stores are visible after it

    // ---- phase B: conv5 (output row q) --------------------------------------------------------------------------------
    {
        const float* src = a3 + (size_t)b * 32 * 16 * 16;
        for (int i = tid; i < 32 * 3 * 4; i += 256) {
            const int c4 = i & 3, t = i >> 2, lr = t % 3, ci = t / 3, row = 2 * q - 1 + lr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0) v = __ldcg(reinterpret_cast<const float4*>(src + ((size_t)ci * 16 + row) * 16) + c4);
            float* d = s_in5 + (ci * 3 + lr) * kTailIn5P + 1 + 4 * c4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        const int ow = wid;   // warp: one pixel of the row; lane: output channel
        float ac[3] = {0.f, 0.f, 0.f};   // one chain per kernel row: three independent FMA chains per thread
#pragma unroll 8
        for (int ci = 0; ci < 32; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* ip = s_in5 + (ci * 3 + kh) * kTailIn5P + 2 * ow;
                const float* wp = s_w5 + (ci * 9 + kh * 3) * kTailWP + lane;
                ac[kh] = fmaf(wp[2 * kTailWP], ip[2], fmaf(wp[kTailWP], ip[1], fmaf(wp[0], ip[0], ac[kh])));
            }
        }
        const float acc = (ac[0] + ac[1]) + ac[2];
        a4[(((size_t)b * 32 + lane) * 8 + q) * 8 + ow] = leaky(acc + __ldg(b5 + lane));
    }
    cluster.sync();   // release/acquire at cluster scope: the peers' a3 / a4 stores are visible after it

    // ---- phase C: fc1 (one output per warp), fc2 in CTA 0 -------------------------------------------------------------
    {
        const float4* f = reinterpret_cast<const float4*>(a4 + (size_t)b * kFc1In);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kFc1In / 128; ++i) {
            const float4 x = __ldcg(f + lane + 32 * i), c = wrow[i];
            acc = fmaf(x.x, c.x, acc); acc = fmaf(x.y, c.y, acc); acc = fmaf(x.z, c.z, acc); acc = fmaf(x.w, c.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float v = leaky(acc + __ldg(b1 + o1));
            s_h[wid] = v;
            h[b * kFc1Out + o1] = v;
        }
    }
    cluster.sync();
    if (q == 0) {
        if (tid < kFc1Out) s_all[tid] = cluster.map_shared_rank(s_h, tid >> 3)[tid & 7];
        __syncthreads();
        if (tid < kFeat) {
            float acc = 0.f;
            for (int o = 0; o < kFc1Out; ++o) acc = fmaf(s_all[o], __ldg(w2 + tid * kFc1Out + o), acc);
            feat[b * kFeat + tid] = acc + __ldg(b2 + tid);
        }
    }
    cluster.sync();  // keep the peers' shared memory alive until CTA 0 has read it
}

}  // namespace dd
