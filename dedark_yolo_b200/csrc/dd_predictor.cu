// dd_predictor.cu -- a4 + a5: bilinear resize to 256x256 and the parameter-predictor CNN, forward and
// backward.  conv2..conv5 (the dense GEMMs) run on the tensor cores (dd_conv_tc.cuh: tcgen05 kind::tf32 with split
// operands, "3xTF32", because the 1e-5 output gate does not survive plain TF32 operand rounding); conv1 (K = 27),
// the resize and the two small FC layers stay on the CUDA cores.
//
// Reference: nn/modules/llie.py:43 (F.interpolate bilinear, align_corners=False),
//            nn/modules/common.py:9-23 (ConvBlock: Conv2d k3 s2 p1 + LeakyReLU 0.1),
//            nn/modules/common.py:52-78 (ExtractParameters2: 5 conv blocks, fc1 2048->64, fc2 64->15).
//
// Layout: every tensor NCHW fp32.  Activations live in the caller's workspace (dd_layout.cuh).  All
// reductions are fixed-order (per-thread serial loops, block_sum, split partials summed in index
// order): bit-reproducible run to run.
#include <cooperative_groups.h>
#include <cstring>

#include "dd_common.cuh"
#include <cuda_bf16.h>

#include "dd_conv_tc.cuh"
#include "dd_conv_tiled.cuh"
#include "dd_layout.cuh"
#include "dd_predictor_tail.cuh"

namespace dd {

// -------------------------------------------------------------------------------------------------
// bilinear source index, ATen upsample_bilinear2d semantics (align_corners=False)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int dst, float scale, int n, int& i0, int& i1, float& lam) {
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i0 = i0 > n - 1 ? n - 1 : i0;
    i1 = i0 < n - 1 ? i0 + 1 : i0;
    lam = s - (float)i0;
}

__device__ __forceinline__ float ld_elem(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_elem(const __nv_bfloat16* p) {
    return __uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}
template <typename T>
__global__ void __launch_bounds__(256)
resize256_kernel(const T* __restrict__ x, float* __restrict__ r, int B, int H, int W) {
    pdl_begin();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // output column (256 per row)
    const int i = blockIdx.y;                              // output row
    const int plane = blockIdx.z;                          // b*3 + ch
    if (j >= DD_RESIZE) return;
    const float sh = (float)H / (float)DD_RESIZE, sw = (float)W / (float)DD_RESIZE;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(i, sh, H, y0, y1, ly);
    bilinear_src(j, sw, W, x0, x1, lx);
    const T* p = x + (size_t)plane * H * W;
    const float v00 = ld_elem(p + (size_t)y0 * W + x0), v01 = ld_elem(p + (size_t)y0 * W + x1);
    const float v10 = ld_elem(p + (size_t)y1 * W + x0), v11 = ld_elem(p + (size_t)y1 * W + x1);
    r[((size_t)plane * DD_RESIZE + i) * DD_RESIZE + j] = bilerp(v00, v01, v10, v11, lx, ly);
}

// adjoint as a gather (deterministic): every source pixel sums the output pixels that sampled it
__global__ void __launch_bounds__(256)
resize256_bwd_kernel(const float* __restrict__ dr, float* __restrict__ dx, int B, int H, int W) {
    pdl_begin();
    const int xs = blockIdx.x * blockDim.x + threadIdx.x;
    const int ys = blockIdx.y;
    const int plane = blockIdx.z;
    if (xs >= W) return;
    const float sh = (float)H / (float)DD_RESIZE, sw = (float)W / (float)DD_RESIZE;
    int ilo = (int)floorf(((float)ys - 0.5f) / sh - 0.5f) - 1, ihi = (int)ceilf(((float)ys + 1.5f) / sh - 0.5f) + 1;
    int jlo = (int)floorf(((float)xs - 0.5f) / sw - 0.5f) - 1, jhi = (int)ceilf(((float)xs + 1.5f) / sw - 0.5f) + 1;
    ilo = ilo < 0 ? 0 : ilo; jlo = jlo < 0 ? 0 : jlo;
    ihi = ihi > DD_RESIZE - 1 ? DD_RESIZE - 1 : ihi; jhi = jhi > DD_RESIZE - 1 ? DD_RESIZE - 1 : jhi;
    const float* d = dr + (size_t)plane * DD_RESIZE * DD_RESIZE;
    float acc = 0.f;
    for (int i = ilo; i <= ihi; ++i) {
        int y0, y1; float ly;
        bilinear_src(i, sh, H, y0, y1, ly);
        const float wy = (y0 == ys ? 1.f - ly : 0.f) + (y1 == ys ? ly : 0.f);
        if (wy == 0.f) continue;
        float racc = 0.f;
        for (int j = jlo; j <= jhi; ++j) {
            int x0, x1; float lx;
            bilinear_src(j, sw, W, x0, x1, lx);
            const float wx = (x0 == xs ? 1.f - lx : 0.f) + (x1 == xs ? lx : 0.f);
            racc = fmaf(wx, __ldg(d + (size_t)i * DD_RESIZE + j), racc);
        }
        acc = fmaf(wy, racc, acc);
    }
    dx[((size_t)plane * H + ys) * W + xs] += acc;
}

// -------------------------------------------------------------------------------------------------
// data gradient of the same conv.  thread = one 2x2 quad of INPUT pixels x CI_T input channels; the four
// parities of a stride-2 3x3 conv touch exactly the 2x2 output neighbourhood (a..a+1, c..c+1):
//   (even,even): w11 d(a,c)                    (even,odd): w10 d(a,c+1) + w12 d(a,c)
//   (odd,even):  w01 d(a+1,c) + w21 d(a,c)     (odd,odd):  w00 d(a+1,c+1) + w02 d(a+1,c) + w20 d(a,c+1) + w22 d(a,c)
// The result is multiplied by LeakyReLU'(act_in) so the stored tensor is the gradient w.r.t. the previous
// layer's PRE-activation (act_in == nullptr for the network input).
// -------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int CI_T>
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ w, const float* __restrict__ act_in,
                  float* __restrict__ din, int B) {
    pdl_begin();
    constexpr int HO = HIN / 2;
    constexpr int CIG = CIN / CI_T;
    __shared__ float sw[COUT * 9 * CIN];  // [co][k][ci]
    for (int i = threadIdx.x; i < COUT * 9 * CIN; i += blockDim.x) {  // ci fastest: conflict-free smem writes
        const int ci = i % CIN, t = i / CIN, k = t % 9, co = t / 9;
        sw[i] = __ldg(w + (co * CIN + ci) * 9 + k);
    }
    __syncthreads();
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (long long)B * HO * HO * CIG) return;
    const int q = (int)(item % (HO * HO));
    const int cig = (int)((item / (HO * HO)) % CIG);
    const int b = (int)(item / ((long long)HO * HO * CIG));
    const int a = q / HO, c = q % HO;
    const bool a1 = a + 1 < HO, c1 = c + 1 < HO;

    float acc[4][CI_T];  // quad order: ee, eo, oe, oo
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CI_T; ++j) acc[i][j] = 0.f;

    const float* db = dpre + (size_t)b * COUT * HO * HO;
    for (int co = 0; co < COUT; ++co) {
        const float* dp = db + ((size_t)co * HO + a) * HO + c;
        const float d00 = __ldg(dp);
        const float d01 = c1 ? __ldg(dp + 1) : 0.f;
        const float d10 = a1 ? __ldg(dp + HO) : 0.f;
        const float d11 = (a1 && c1) ? __ldg(dp + HO + 1) : 0.f;
        const float* wp = sw + co * 9 * CIN + cig * CI_T;
#pragma unroll
        for (int j = 0; j < CI_T; ++j) {
            const float w00 = wp[0 * CIN + j], w01 = wp[1 * CIN + j], w02 = wp[2 * CIN + j];
            const float w10 = wp[3 * CIN + j], w11 = wp[4 * CIN + j], w12 = wp[5 * CIN + j];
            const float w20 = wp[6 * CIN + j], w21 = wp[7 * CIN + j], w22 = wp[8 * CIN + j];
            acc[0][j] = fmaf(w11, d00, acc[0][j]);
            acc[1][j] = fmaf(w10, d01, fmaf(w12, d00, acc[1][j]));
            acc[2][j] = fmaf(w01, d10, fmaf(w21, d00, acc[2][j]));
            acc[3][j] = fmaf(w00, d11, fmaf(w02, d10, fmaf(w20, d01, fmaf(w22, d00, acc[3][j]))));
        }
    }
#pragma unroll
    for (int j = 0; j < CI_T; ++j) {
        const int ci = cig * CI_T + j;
        const size_t base = (((size_t)b * CIN + ci) * HIN + 2 * a) * HIN + 2 * c;
        float2 top = make_float2(acc[0][j], acc[1][j]), bot = make_float2(acc[2][j], acc[3][j]);
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

// -------------------------------------------------------------------------------------------------
// fully connected layers
// -------------------------------------------------------------------------------------------------
// fc1 + fc2 of one image per CLUSTER of 4 CTAs: CTA r computes h[b][16r .. 16r+15] = leaky(flat[b] . W1[o] + b1[o])
// (warp w: outputs 16r+2w, 16r+2w+1; all weight loads of a lane in flight at once, requested BEFORE the grid dependency is
// resolved -- the weights are parameters, no kernel of this library writes them), every CTA stores its quarter straight into
// CTA 0's shared memory (distributed shared memory), and CTA 0 finishes feat[b][j] = h[b] . W2[j] + b2[j] with 16 lanes per
// output (W2 staged in shared memory before the wait as well).  In the stream this kernel is pure latency: the serial
// 64-load fc2 loop and a second cluster barrier of the first version cost 5 of its 9 us.
constexpr int kFcCluster = 4;
// A COHERENT ld.global in a volatile asm stays where it is written relative to griddepcontrol.wait.  ld.global.nc does not: ptxas
// reschedules it freely in BOTH directions (__ldg builtin and inline PTX alike).  When this kernel first read its weights
// before the wait and the producer's data behind it, both with __ldg, the producer's loads were hoisted ABOVE the wait -- a
// stale read.  Rule (checked on the SASS of every kernel by tests/test_sass_pdl.py): no global access in front of ACQBULK,
// except here, where weights are read with ldg_pinned and the producer's data with __ldcg (ld.global.cg, never hoisted).
__device__ __forceinline__ float4 ldg_pinned(const float4* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(256)
fc_fwd_kernel(const float* __restrict__ flat, const float* __restrict__ w1, const float* __restrict__ b1,
              const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ h, float* __restrict__ feat) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int PER_CTA = kFc1Out / kFcCluster;  // 16
    constexpr int NV = kFc1In / 128;               // float4 per lane and row
    __shared__ __align__(16) float s_all[kFc1Out];
    __shared__ __align__(16) float s_w2[kFeat * kFc1Out];
    const int b = blockIdx.x / kFcCluster, r = (int)cluster.block_rank(), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int o0 = r * PER_CTA + 2 * wid;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // first half of a split cluster barrier: CTA 0 must have started before a peer stores into its shared memory; the wait
    // (below, in front of those stores) is free by then
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    float4 c0[NV], c1[NV];
    {
        const float4* wa = reinterpret_cast<const float4*>(w1 + (size_t)o0 * kFc1In);
        const float4* wb = wa + kFc1In / 4;
#pragma unroll
        for (int i = 0; i < NV; ++i) { c0[i] = ldg_pinned(wa + lane + 32 * i); c1[i] = ldg_pinned(wb + lane + 32 * i); }
    }
    const float bias0 = __ldg(b1 + o0), bias1 = __ldg(b1 + o0 + 1);
    if (r == 0 && threadIdx.x < kFeat * kFc1Out / 4)
        reinterpret_cast<float4*>(s_w2)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(w2) + threadIdx.x);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const float4* f = reinterpret_cast<const float4*>(flat + (size_t)b * kFc1In);
    float4 x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) x[i] = __ldcg(f + lane + 32 * i);   // NOT __ldg: the compiler hoists ld.global.nc above the wait
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        acc0 = fmaf(x[i].x, c0[i].x, acc0); acc0 = fmaf(x[i].y, c0[i].y, acc0); acc0 = fmaf(x[i].z, c0[i].z, acc0); acc0 = fmaf(x[i].w, c0[i].w, acc0);
        acc1 = fmaf(x[i].x, c1[i].x, acc1); acc1 = fmaf(x[i].y, c1[i].y, acc1); acc1 = fmaf(x[i].z, c1[i].z, acc1); acc1 = fmaf(x[i].w, c1[i].w, acc1);
    }
    acc0 = warp_sum(acc0);
    acc1 = warp_sum(acc1);
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    if (lane == 0) {
        const float v0 = leaky(acc0 + bias0), v1 = leaky(acc1 + bias1);
        float* dst = cluster.map_shared_rank(s_all, 0);   // CTA 0's copy
        dst[o0] = v0;
        dst[o0 + 1] = v1;
        h[b * kFc1Out + o0] = v0;
        h[b * kFc1Out + o0 + 1] = v1;
    }
    cluster.sync();   // the only barrier: the peers' stores into CTA 0 are complete and visible; nobody reads the peers' memory
    if (r == 0) {
        const int j = threadIdx.x >> 4, p = threadIdx.x & 15;   // 16 lanes per output, 4 products each, xor-tree in a fixed order
        const int jj = j < kFeat ? j : kFeat - 1;               // (the 16th group only keeps the shuffles full-warp)
        const float4 hv = reinterpret_cast<const float4*>(s_all)[p], wv = reinterpret_cast<const float4*>(s_w2)[jj * 16 + p];
        float acc = fmaf(hv.w, wv.w, fmaf(hv.z, wv.z, fmaf(hv.y, wv.y, hv.x * wv.x)));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (p == 0 && j < kFeat) feat[b * kFeat + j] = acc + __ldg(b2 + j);
    }
}

// The whole FC backward in ONE launch.  dhpre[b][o] = leaky'(h[b][o]) * sum_j dfeat[b][j] W2[j][o] is so cheap (15 MACs)
// that every CTA recomputes the few values it needs in shared memory instead of waiting for a separate kernel:
//   CTAs [0, n_w)         dW1[o][i] = sum_b dhpre[b][o] flat[b][i]  (256 consecutive i of ONE o per CTA), db1[o]
//   CTAs [n_w, n_w + n_d) dpre5[b][i] = leaky'(a5[b][i]) * sum_o dhpre[b][o] W1[o][i]  (256 consecutive i of ONE image)
//   last 4 CTAs           dW2[j][o] = sum_b dfeat[b][j] h[b][o] (one element per thread),  db2[j] = sum_b dfeat[b][j]
constexpr int kFc2Ctas = (kFeat * kFc1Out + 255) / 256;   // 4
__global__ void __launch_bounds__(256)
fc_bwd_kernel(const float* __restrict__ dfeat, const float* __restrict__ h, const float* __restrict__ w2,
              const float* __restrict__ flat, const float* __restrict__ w1, float* __restrict__ dw2, float* __restrict__ db2,
              float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dpre5, int B, int n_w, int n_d,
              const PushCtx px) {
    pdl_wait();     // wait, then launch: the tensor-core kernel behind this one copies its (forward-pass) weights ahead of its own wait
    pdl_launch();
    static_assert(kFc1In % 256 == 0, "a CTA never straddles two rows of W1");
    const unsigned int tag = push_tag(px);
    __shared__ float s_dh[256];
    const int tid = threadIdx.x, bid = blockIdx.x;
    if (bid < n_w) {
        const int t = bid * 256 + tid, o = t / kFc1In, i = t % kFc1In;
        float acc = 0.f, bsum = 0.f;
        for (int b0 = 0; b0 < B; b0 += 256) {
            const int nb = min(256, B - b0);
            if (tid < nb) {
                const int b = b0 + tid;
                float v = 0.f;
#pragma unroll
                for (int j = 0; j < kFeat; ++j) v = fmaf(__ldg(dfeat + b * kFeat + j), __ldg(w2 + j * kFc1Out + o), v);
                s_dh[tid] = leaky_grad(__ldg(h + b * kFc1Out + o), v);
            }
            __syncthreads();
            for (int bb = 0; bb < nb; ++bb) {
                acc = fmaf(s_dh[bb], __ldg(flat + (size_t)(b0 + bb) * kFc1In + i), acc);
                bsum += s_dh[bb];
            }
            __syncthreads();
        }
        dw1[t] = acc;
        if (i == 0) db1[o] = bsum;
        if (px.world > 1) {
            push_grad(px, tag, kGradOffFc1W + t, acc);
            if (i == 0) push_grad(px, tag, kGradOffFc1B + o, bsum);
        }
    } else if (bid < n_w + n_d) {
        const int t = (bid - n_w) * 256 + tid, b = t / kFc1In, i = t % kFc1In;  // b is the same for the whole CTA
        if (tid < kFc1Out && b < B) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < kFeat; ++j) v = fmaf(__ldg(dfeat + b * kFeat + j), __ldg(w2 + j * kFc1Out + tid), v);
            s_dh[tid] = leaky_grad(__ldg(h + b * kFc1Out + tid), v);
        }
        __syncthreads();
        if (b >= B) return;
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < kFc1Out; ++o) acc = fmaf(s_dh[o], __ldg(w1 + (size_t)o * kFc1In + i), acc);
        dpre5[t] = leaky_grad(flat[t], acc);
    } else {   // the last kFc2Ctas CTAs: one dW2 element per thread (one CTA looping four times was the longest chain of the kernel)
        const int t = (bid - n_w - n_d) * 256 + tid;
        if (t < kFeat * kFc1Out) {
            const int j = t / kFc1Out, o = t % kFc1Out;
            float acc = 0.f;
            for (int b = 0; b < B; ++b) acc = fmaf(__ldg(dfeat + b * kFeat + j), __ldg(h + b * kFc1Out + o), acc);
            dw2[t] = acc;
            if (px.world > 1) push_grad(px, tag, kGradOffFc2W + t, acc);
        }
        if (bid == n_w + n_d && tid < kFeat) {
            float acc = 0.f;
            for (int b = 0; b < B; ++b) acc += __ldg(dfeat + b * kFeat + tid);
            db2[tid] = acc;
            if (px.world > 1) push_grad(px, tag, kGradOffFc2B + tid, acc);
        }

    }
}

// -------------------------------------------------------------------------------------------------
// host-side launch helpers
// -------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int HIN, int CI_T>
static void launch_conv_dgrad(const float* dpre, const float* w, const float* act_in, float* din, int B,
                              cudaStream_t st) {
    constexpr int HO = HIN / 2;
    const long long items = (long long)B * HO * HO * (CIN / CI_T);
    launch_pdl(conv_dgrad_kernel<CIN, COUT, HIN, CI_T>, dim3((unsigned)((items + 255) / 256)), dim3(256), 0, st, dpre, w, act_in, din, B);
    count_launch();
}

// ---- conv1: persistent CUDA-core kernel with TMA-staged tiles (dd_conv_tiled.cuh).  The same launch carries, in extra
// CTAs, the preparation of the tensor-core weights of conv2..conv5 (independent of conv1; one launch less on the
// critical path).
__global__ void __launch_bounds__(256)
conv1_fwd_prep_kernel(const __grid_constant__ CUtensorMap rmap, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int ntiles, int n_conv, const tc::PrepJobs jobs) {
    pdl_launch();
    extern __shared__ __align__(128) unsigned char smem_c1f[];
    __shared__ __align__(8) uint64_t full[2];
    if ((int)blockIdx.x < n_conv) {
        if (threadIdx.x == 0) {
            mbar_init(&full[0], 1);
            mbar_init(&full[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        pdl_wait();
        conv1_fwd_body(blockIdx.x, n_conv, &rmap, w, bias, out, ntiles, smem_c1f, full);
    } else {
        pdl_wait();
        tc::prep_weights_body(jobs, (int)blockIdx.x - n_conv);
    }
}

static int launch_conv1_prep(const float* in, const float* w, const float* b, float* out, const tc::PrepJobs& jobs, int B, cudaStream_t st) {
    constexpr size_t smem = conv1_fwd_smem();
    DD_ENSURE_SMEM(conv1_fwd_prep_kernel, smem, "conv1_fwd_prep_kernel");
    CUtensorMap rmap;
    if (!make_tensor_map_3d(&rmap, in, B * kC1In, kC1HIn, kC1HIn, kC1BoxW, 2 * kC1FwdTH + 1, kC1In)) {
        set_error("dd_predictor_fwd: cannot encode the TMA tensor map of the resized batch (r must be 16-byte aligned)");
        return DD_ERR_CUDA;
    }
    const int ntiles = B * (kC1HOut / kC1FwdTH) * (kC1HOut / kC1TW);
    const int n_conv = ntiles < 2 * sm_count() ? ntiles : 2 * sm_count();
    launch_pdl(conv1_fwd_prep_kernel, dim3(n_conv + 4 * tc::kPrepBlocksPerJob), dim3(256), smem, st, rmap, w, b, out, ntiles, n_conv, jobs);
    count_launch();
    return DD_OK;
}

// ---- conv2..conv5: tensor-core kernels (dd_conv_tc.cuh) --------------------------------------------
template <int CIN, int COUT, int HIN>
static int launch_tc_fwd(const float* in, const float* wprep, const float* bias, float* out, int B, cudaStream_t st) {
    constexpr int HO = HIN / 2;
    constexpr size_t smem = tc::conv_tc_fwd_smem<CIN, COUT>();
    auto kern = tc::conv_tc_fwd<CIN, COUT, HIN>;
    DD_ENSURE_SMEM(kern, smem, "conv_tc_fwd");
    const int total = B * HO * HO, ntiles = (total + 127) / 128;
    launch_pdl(kern, dim3(ntiles < sm_count() ? ntiles : sm_count()), dim3(288), smem, st, in, wprep, bias, out, total);  // persistent: <= 1 CTA per SM
    count_launch();
    return DD_OK;
}

// weight gradient (one slice per CTA -> `partial`) and data gradient of one layer in one launch of persistent CTAs; the
// SMs are divided between the two roles in proportion to their estimated MMA time.  Returns the slice count in *n_slices.
template <int CIN, int COUT, int HIN>
static int launch_tc_bwd(const float* in, const float* dpre, const float* wprep_dgrad, const float* act_in, float* partial,
                         float* din, int* n_slices, int B, cudaStream_t st) {
    constexpr int HO = HIN / 2;
    constexpr size_t smem = tc::conv_tc_bwd_smem<CIN, COUT, HIN>();
    auto kern = tc::conv_tc_bwd<CIN, COUT, HIN>;
    DD_ENSURE_SMEM(kern, smem, "conv_tc_bwd");
    const int total = B * HO * HO;
    const int wtiles = (total + tc::WgradCfg<CIN>::PXT - 1) / tc::WgradCfg<CIN>::PXT, dtiles = (total + 127) / 128;
    const int ctas = sm_count() < kTcMaxCtas ? sm_count() : kTcMaxCtas;
    const double cw = (double)wtiles * 2500.0, cd = (double)dtiles * (CIN == 32 ? 7500.0 : 4500.0);  // measured cycles per tile (profiles/microbench/tc_probe.cu)
    int n_w = (int)(ctas * cw / (cw + cd) + 0.5);
    n_w = n_w < 1 ? 1 : (n_w > ctas - 1 ? ctas - 1 : n_w);
    n_w = n_w > wtiles ? wtiles : n_w;
    const int n_d = dtiles < ctas - n_w ? dtiles : ctas - n_w;
    launch_pdl(kern, dim3(n_w + n_d), dim3(288), smem, st, in, dpre, wprep_dgrad, act_in, partial, din, n_w, total);
    count_launch();
    *n_slices = n_w;
    return DD_OK;
}

static bool tensors_ok(const dd_predictor_tensors* t) {
    if (!t) return false;
    for (int i = 0; i < 5; ++i)
        if (!t->conv_w[i] || !t->conv_b[i]) return false;
    return t->fc1_w && t->fc1_b && t->fc2_w && t->fc2_b;
}

}  // namespace dd

extern "C" int dd_resize256(const float* x, float* r, int B, int H, int W, void* stream_) {
    using namespace dd;
    DD_REQUIRE(x && r && B > 0 && H > 0 && W > 0, DD_ERR_INVALID, "dd_resize256: bad arguments");
    DD_REQUIRE((long long)B * 3 <= 65535, DD_ERR_INVALID, "dd_resize256: B too large (%d)", B);
    dim3 grid(1, DD_RESIZE, B * 3);
    launch_pdl(resize256_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream_, x, r, B, H, W);
    count_launch();
    return check_launch("dd_resize256");
}

extern "C" int dd_resize256_ex(const void* x, int x_dtype, float* r, int B, int H, int W, void* stream_) {
    using namespace dd;
    if (x_dtype == DD_F32) return dd_resize256(reinterpret_cast<const float*>(x), r, B, H, W, stream_);
    DD_REQUIRE(x_dtype == DD_BF16, DD_ERR_INVALID, "dd_resize256_ex: unknown dtype %d", x_dtype);
    DD_REQUIRE(x && r && B > 0 && H > 0 && W > 0, DD_ERR_INVALID, "dd_resize256_ex: bad arguments");
    DD_REQUIRE((long long)B * 3 <= 65535, DD_ERR_INVALID, "dd_resize256_ex: B too large (%d)", B);
    dim3 grid(1, DD_RESIZE, B * 3);
    launch_pdl(resize256_kernel<__nv_bfloat16>, grid, dim3(256), 0, (cudaStream_t)stream_, reinterpret_cast<const __nv_bfloat16*>(x), r, B, H, W);
    count_launch();
    return check_launch("dd_resize256_ex");
}

extern "C" int dd_resize256_bwd(const float* dr, float* dx, int B, int H, int W, void* stream_) {
    using namespace dd;
    DD_REQUIRE(dr && dx && B > 0 && H > 0 && W > 0, DD_ERR_INVALID, "dd_resize256_bwd: bad arguments");
    DD_REQUIRE((long long)B * 3 <= 65535 && H <= 65535, DD_ERR_INVALID, "dd_resize256_bwd: shape too large");
    dim3 grid((W + 255) / 256, H, B * 3);
    launch_pdl(resize256_bwd_kernel, grid, dim3(256), 0, (cudaStream_t)stream_, dr, dx, B, H, W);
    count_launch();
    return check_launch("dd_resize256_bwd");
}

extern "C" int dd_predictor_fwd(const float* r, const dd_predictor_tensors* w, float* acts, float* feat, int B,
                                void* stream_) {
    using namespace dd;
    cudaStream_t st = (cudaStream_t)stream_;
    DD_REQUIRE(r && acts && feat && B > 0 && tensors_ok(w), DD_ERR_INVALID, "dd_predictor_fwd: bad arguments");
    float* a[6];
    for (int l = 0; l < 6; ++l) a[l] = acts + pred_act_offset(l, B);
    float* prep = acts + predictor_acts_elems(B);
    {
        tc::PrepJobs jobs;
        for (int l = 1; l < 5; ++l)
            jobs.j[l - 1] = tc::PrepJob{w->conv_w[l], prep + pred_prep_offset(l), prep + pred_prep_offset(l) + pred_prep_fwd_elems(l),
                                        pred_cin(l), pred_cout(l)};
        if (int e = launch_conv1_prep(r, w->conv_w[0], w->conv_b[0], a[0], jobs, B, st)) return e;
    }
    if (int e = launch_tc_fwd<16, 32, 128>(a[0], prep + pred_prep_offset(1), w->conv_b[1], a[1], B, st)) return e;
    if (int e = launch_tc_fwd<32, 32, 64>(a[1], prep + pred_prep_offset(2), w->conv_b[2], a[2], B, st)) return e;
    if (tail_fused()) {  // experiment: conv4, conv5, fc1, fc2 of one image per cluster of 8 CTAs (dd_predictor_tail.cuh)
        static const cudaError_t attr = cudaFuncSetAttribute(predictor_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTailSmem);
        DD_REQUIRE(attr == cudaSuccess, DD_ERR_CUDA, "dd_predictor_fwd: shared memory opt-in failed: %s", cudaGetErrorString(attr));
        launch_pdl_cluster(predictor_tail_kernel, dim3(B * kTailCluster), dim3(256), kTailSmem, st, kTailCluster, (const float*)a[2],
                           (const float*)w->conv_w[3], (const float*)w->conv_b[3], (const float*)w->conv_w[4], (const float*)w->conv_b[4],
                           (const float*)w->fc1_w, (const float*)w->fc1_b, (const float*)w->fc2_w, (const float*)w->fc2_b, a[3], a[4], a[5], feat);
        count_launch();
        return check_launch("dd_predictor_fwd");
    }
    if (int e = launch_tc_fwd<32, 32, 32>(a[2], prep + pred_prep_offset(3), w->conv_b[3], a[3], B, st)) return e;
    if (int e = launch_tc_fwd<32, 32, 16>(a[3], prep + pred_prep_offset(4), w->conv_b[4], a[4], B, st)) return e;
    launch_pdl_cluster(fc_fwd_kernel, dim3(B * kFcCluster), dim3(256), 0, st, kFcCluster, (const float*)a[4], (const float*)w->fc1_w, (const float*)w->fc1_b,
               (const float*)w->fc2_w, (const float*)w->fc2_b, a[5], feat);
    count_launch();
    return check_launch("dd_predictor_fwd");
}

namespace dd {
// part: 0 = the whole backward; 1 = the fully connected layers only (their gradients -- 80 % of the bytes -- are final after this
// one launch, so a caller can start reducing them over the ranks while the convolutions run); 2 = the convolutions only
static int predictor_bwd_impl(const float* r, const dd_predictor_tensors* w, const float* acts, const float* dfeat,
                              const dd_predictor_tensors* g, float* dr, int B, void* ws, size_t ws_bytes,
                              const dd_peer_exchange* pxh, cudaStream_t st, int part = 0) {
    PushCtx px;
    memset(&px, 0, sizeof(px));
    if (pxh && pxh->world > 1) {
        px.rank = pxh->rank;
        px.world = pxh->world;
        for (int i = 0; i < pxh->world; ++i) px.buf[i] = reinterpret_cast<unsigned char*>(pxh->buf[i]);
    }
    DD_REQUIRE(r && acts && dfeat && B > 0 && tensors_ok(w) && tensors_ok(g), DD_ERR_INVALID,
               "dd_predictor_bwd: bad arguments");
    DD_REQUIRE(ws && ws_bytes >= predictor_bwd_ws_bytes(B), DD_ERR_WORKSPACE, "dd_predictor_bwd: workspace %zu < %zu",
               ws_bytes, predictor_bwd_ws_bytes(B));
    const float* a[6];
    float* d[6];
    for (int l = 0; l < 6; ++l) {
        a[l] = acts + pred_act_offset(l, B);
        d[l] = reinterpret_cast<float*>(ws) + pred_act_offset(l, B);
    }
    float* partial = reinterpret_cast<float*>(ws) + predictor_acts_elems(B);

    if (part != 2) {
        const int n_w = kFc1Out * kFc1In / 256, n_d = (B * kFc1In + 255) / 256;
        launch_pdl(fc_bwd_kernel, dim3(n_w + n_d + kFc2Ctas), dim3(256), 0, st, dfeat, a[5], (const float*)w->fc2_w, a[4], (const float*)w->fc1_w,
                   g->fc2_w, g->fc2_b, g->fc1_w, g->fc1_b, d[4], B, n_w, n_d, px);
        count_launch();
        if (part == 1) return check_launch("dd_predictor_bwd_part(fc)");
    }
    // conv5 .. conv2: weight-gradient slices and the data gradient of a layer in one tensor-core launch; conv1: weight
    // gradient on the CUDA cores; one deferred reduction of all slice buffers at the end
    const float* prep = acts + predictor_acts_elems(B);
    float* pl[5];
    int nsl[5];
    for (int l = 0; l < 5; ++l) pl[l] = partial + pred_wgrad_partial_offset(l, B);
    nsl[0] = B * kWgradC1Slices < kTcMaxCtas ? B * kWgradC1Slices : kTcMaxCtas;  // persistent: one slice per CTA
    auto dg = [&](int l) { return prep + pred_prep_offset(l) + pred_prep_fwd_elems(l); };
    if (int e = launch_tc_bwd<32, 32, 16>(a[3], d[4], dg(4), a[3], pl[4], d[3], &nsl[4], B, st)) return e;
    if (int e = launch_tc_bwd<32, 32, 32>(a[2], d[3], dg(3), a[2], pl[3], d[2], &nsl[3], B, st)) return e;
    if (int e = launch_tc_bwd<32, 32, 64>(a[1], d[2], dg(2), a[1], pl[2], d[1], &nsl[2], B, st)) return e;
    if (int e = launch_tc_bwd<16, 32, 128>(a[0], d[1], dg(1), a[0], pl[1], d[0], &nsl[1], B, st)) return e;
    {   // first layer (CIN = 3): persistent CUDA-core kernel, tiles of r and of d(conv1) by TMA
        constexpr size_t smem = conv1_wgrad_smem();
        DD_ENSURE_SMEM(conv1_wgrad_kernel, smem, "conv1_wgrad_kernel");
        CUtensorMap rmap, dmap;
        if (!make_tensor_map_3d(&rmap, r, B * kC1In, kC1HIn, kC1HIn, kC1BoxW, 2 * kC1WgTH + 1, kC1In) ||
            !make_tensor_map_3d(&dmap, d[0], B * kC1Out, kC1HOut, kC1HOut, kC1TW, kC1WgTH, kC1Out)) {
            set_error("dd_predictor_bwd: cannot encode the TMA tensor maps (r and the workspace must be 16-byte aligned)");
            return DD_ERR_CUDA;
        }
        launch_pdl(conv1_wgrad_kernel, dim3(nsl[0]), dim3(256), smem, st, rmap, dmap, pl[0], B * kWgradC1Slices);
    }
    {
        tc::ReduceJobs jobs;
        int block0 = 0;
        for (int l = 0; l < 5; ++l) {
            const int n = l == 0 ? 432 + 16 : (9 * pred_cin(l) + 1) * 32;
            jobs.j[l] = tc::ReduceJob{pl[l], g->conv_w[l], g->conv_b[l], nsl[l], n, l == 0 ? n : pred_wgrad_rp(l) * 32,
                                      l == 0 ? 0 : pred_cin(l), 432, block0, grad_off_conv_w(l), grad_off_conv_b(l)};
            block0 += (n + tc::kReduceOut - 1) / tc::kReduceOut;
        }
        launch_pdl(tc::wgrad_reduce_kernel, dim3(block0), dim3(tc::kReduceWarps * 32), 0, st, jobs, px);
    }
    count_launch(2);
    if (px.world > 1) {
        tc::GradPtrs gp;
        for (int l = 0; l < 5; ++l) {
            gp.p[2 * l] = g->conv_w[l];
            gp.p[2 * l + 1] = g->conv_b[l];
        }
        gp.p[10] = g->fc1_w; gp.p[11] = g->fc1_b; gp.p[12] = g->fc2_w; gp.p[13] = g->fc2_b;
        // 64 CTAs, not one per SM: the kernel waits for other ranks; leaving SMs free keeps a second rank on the SAME device
        // (the single-GPU test, MPS-style sharing) schedulable whatever the register footprints are
        launch_pdl(tc::allreduce_exchange_kernel, dim3(64), dim3(512), 0, st, px, gp);
        count_launch();
    }
    if (dr) launch_conv_dgrad<3, 16, 256, 3>(d[0], w->conv_w[0], nullptr, dr, B, st);
    return check_launch("dd_predictor_bwd");
}
}  // namespace dd

extern "C" int dd_predictor_bwd(const float* r, const dd_predictor_tensors* w, const float* acts,
                                const float* dfeat, const dd_predictor_tensors* g, float* dr, int B, void* ws,
                                size_t ws_bytes, void* stream_) {
    return dd::predictor_bwd_impl(r, w, acts, dfeat, g, dr, B, ws, ws_bytes, nullptr, (cudaStream_t)stream_);
}

extern "C" int dd_predictor_bwd_part(const float* r, const dd_predictor_tensors* w, const float* acts, const float* dfeat,
                                     const dd_predictor_tensors* g, float* dr, int B, void* ws, size_t ws_bytes, int part,
                                     void* stream_) {
    DD_REQUIRE(part >= 0 && part <= 2, DD_ERR_INVALID, "dd_predictor_bwd_part: part must be 0, 1 or 2 (got %d)", part);
    return dd::predictor_bwd_impl(r, w, acts, dfeat, g, dr, B, ws, ws_bytes, nullptr, (cudaStream_t)stream_, part);
}

extern "C" size_t dd_exchange_bytes(void) { return dd::exchange_bytes(); }

extern "C" int dd_predictor_bwd_allreduce(const float* r, const dd_predictor_tensors* w, const float* acts,
                                          const float* dfeat, const dd_predictor_tensors* g, int B, void* ws,
                                          size_t ws_bytes, const dd_peer_exchange* px, void* stream_) {
    using namespace dd;
    DD_REQUIRE(px && px->world >= 1 && px->world <= DD_MAX_PEERS && px->rank >= 0 && px->rank < px->world, DD_ERR_INVALID,
               "dd_predictor_bwd_allreduce: bad peer exchange descriptor");
    for (int i = 0; i < px->world; ++i)
        DD_REQUIRE(px->buf[i] != nullptr, DD_ERR_INVALID, "dd_predictor_bwd_allreduce: peer buffer %d is null", i);
    return predictor_bwd_impl(r, w, acts, dfeat, g, nullptr, B, ws, ws_bytes, px, (cudaStream_t)stream_);
}
