// dd_common.cuh -- shared helpers for the sm_100a kernels of libdedark_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dedark_b200.h"

namespace dd {

// ---- frozen constants of the reference (filter_cfg.py:17-44, filtersB.py:156,163,247) ------------
constexpr int kFeat = DD_NUM_FEATURES;
constexpr int kSlotDedark = 0, kSlotWb = 1, kSlotGamma = 4, kSlotContrast = 13, kSlotUsm = 14;
constexpr float kLumR = 0.27f, kLumG = 0.67f, kLumB = 0.06f;
constexpr float kWbEps = 1e-5f, kContrastEps = 1e-6f, kTxMin = 0.01f, kGammaClamp = 1e-4f;
constexpr float kDefaultA = 0.8f, kDefaultIcA = 0.5f;
constexpr float kLeaky = 0.1f;
constexpr int kRadius = 12;
constexpr int kTaps = 25;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kLn3 = 1.09861228866810969140f;

// k1[i] = exp(-0.5 ((i-12)/5)^2) / sum, evaluated in fp32 exactly as filtersB.py:158-160 does
// (values recorded from torch; checked by tests/test_oracle.py::test_blur_dense_equals_separable_fp64
// and by the GPU parity tests).  Index = |offset|.
#define DD_K0 0.08077993243932724f  // 0x1.4adfe60000000p-4
#define DD_K1 0.07918038219213486f  // 0x1.4452a60000000p-4
#define DD_K2 0.07456927001476288f  // 0x1.316f8c0000000p-4
#define DD_K3 0.06747306883335114f  // 0x1.145ea40000000p-4
#define DD_K4 0.05865826830267906f  // 0x1.e0874e0000000p-5
#define DD_K5 0.04899550601840019f  // 0x1.915f060000000p-5
#define DD_K6 0.03931981325149536f  // 0x1.421ba00000000p-5
#define DD_K7 0.030317604541778564f  // 0x1.f0b9400000000p-6
#define DD_K8 0.022459832951426506f  // 0x1.6ffb5e0000000p-6
#define DD_K9 0.01598624512553215f  // 0x1.05eb2c0000000p-6
#define DD_K10 0.010932374745607376f  // 0x1.663b680000000p-7
#define DD_K11 0.007183081936091185f  // 0x1.d6c01e0000000p-8
#define DD_K12 0.004534561652690172f  // 0x1.292d520000000p-8

__host__ __device__ __forceinline__ constexpr float tap(int a) {  // a = |offset| in 0..12
    return a == 0 ? DD_K0 : a == 1 ? DD_K1 : a == 2 ? DD_K2 : a == 3 ? DD_K3 : a == 4 ? DD_K4
         : a == 5 ? DD_K5 : a == 6 ? DD_K6 : a == 7 ? DD_K7 : a == 8 ? DD_K8 : a == 9 ? DD_K9
         : a == 10 ? DD_K10 : a == 11 ? DD_K11 : a == 12 ? DD_K12 : 0.0f;
}

// ---- host-side error plumbing --------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(unsigned n = 1);
int check_launch(const char* what);  // cudaGetLastError -> DD_OK / DD_ERR_CUDA

#define DD_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            dd::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

// opt a kernel into > 48 KB of dynamic shared memory once per (call site, device): the attribute call costs a few
// microseconds of host time, which matters for eager launches (the static lives in the enclosing template instance)
#define DD_ENSURE_SMEM(kern, bytes, what)                                                                   \
    do {                                                                                                    \
        static unsigned long long dd_mask__ = 0ull;                                                         \
        int dd_dev__ = 0;                                                                                   \
        cudaGetDevice(&dd_dev__);                                                                           \
        if (!((dd_mask__ >> (dd_dev__ & 63)) & 1ull)) {                                                     \
            cudaError_t dd_e__ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (dd_e__ != cudaSuccess) {                                                                    \
                dd::set_error("cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(dd_e__));            \
                return DD_ERR_CUDA;                                                                         \
            }                                                                                               \
            dd_mask__ |= 1ull << (dd_dev__ & 63);                                                           \
        }                                                                                                   \
    } while (0)

inline int sm_count() {  // SMs of the current device (all GPUs of a node are the same part)
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute: the next kernel in the stream may
// be scheduled as soon as all CTAs of this one have executed griddepcontrol.launch_dependents (or exited); its CTAs run until
// their own griddepcontrol.wait, which returns when this grid has completed and its memory is visible.  That hides the launch
// latency, the prologues and the first loads of the ~20 dependent kernels of a step.  The rules that keep it correct:
//   1. Default (pdl_begin): release, then wait, first thing in the kernel -- nothing touches global memory before the wait.  A grid
//      cannot complete before its predecessors (its working CTAs waited for them).
//   2. Prologues that touch no global memory (barrier initialisation, TMEM allocation, the first __syncthreads) may sit in front
//      of the wait: pdl_launch(); prologue; pdl_wait().  Kernels that allocate TMEM release their dependents AFTER the
//      allocation: a dependent CTA sharing the SM must not take the columns first (it would hold them while waiting for this
//      grid, which waits for the columns).
//   3. A kernel may READ ahead of its wait only what its predecessor cannot be writing:
//      - parameters (no kernel of the library writes them);
//      - image-sized inputs of the step (x, IcA, A, g): the kernels that PRODUCE such tensors (synthesis, dark-channel prior)
//        never release their dependents early (pdl_wait_only), so whatever runs behind them sees them complete;
//      - data written two or more kernels earlier, IF the predecessor waits before it releases (pdl_wait(); pdl_launch();):
//        a dependent CTA then only runs once the predecessor is past its own wait, i.e. once everything older has completed.
//        (When only some CTAs of the predecessor do so, that is enough: the dependent grid starts when EVERY CTA has released.)
//   4. Nothing is WRITTEN to global memory ahead of the wait, except by CTAs whose inputs and outputs are covered by 3 and that
//      never wait at all (the slice reduction of the tensor-core layers, dd_conv_tc.cuh).
//   5. ptxas moves ld.global.nc (__ldg, const __restrict__ loads) across griddepcontrol.wait in BOTH directions.  Data of the
//      predecessor that is read right behind a wait uses ld.global.cg (__ldcg, a volatile asm); loads that must stay in front of
//      it use a coherent ld.global in a volatile asm (ldg_pinned, dd_predictor.cu) or cp.async / TMA.
// tests/test_sass_pdl.py checks the machine code of every kernel: no global access in front of ACQBULK except the listed ones.
__device__ __forceinline__ void pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_only() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool no_pdl = [] { const char* v = getenv("DEDARK_NOPDL"); return v && v[0] == '1'; }();  // debugging aid
    cfg.numAttrs = no_pdl ? 0 : 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    // DEDARK_SYNC=1 (debugging aid): synchronise after every launch and name the first one that fails
    static const bool dbg = [] { const char* v = getenv("DEDARK_SYNC"); return v && v[0] == '1'; }();
    if (dbg) {
        static int seq = 0;
        ++seq;
        const cudaError_t se = cudaDeviceSynchronize();
        if (e != cudaSuccess || se != cudaSuccess)
            fprintf(stderr, "[dedark] launch #%d of this call site (grid %u, block %u, smem %zu) failed: %s / %s\n", seq, grid.x, block.x, smem,
                    cudaGetErrorString(e), cudaGetErrorString(se));
        else
            fprintf(stderr, "[dedark] ok: grid %u block %u smem %zu\n", grid.x, block.x, smem);
    }
    return e;
}

// same, as thread-block clusters of `cluster_x` CTAs (distributed shared memory between the CTAs of a cluster)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x,
                                      Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- peer exchange (dd_predictor_bwd_allreduce) ----------------------------------------------------
// world <= 1: no exchange.  Otherwise gradient element `off` (canonical flat order) is also stored into slot [rank] of
// the current parity in every rank's exchange buffer (peer memory over NVLink) as ONE 8-byte word {value, tag} (one b64 store): the tag
// (epoch + 1) travels with the data, so the receiver polls the element itself -- no fence, no separate flag, no second
// NVLink round trip (the "LL" idea of collective libraries).
struct PushCtx {
    int rank, world;
    unsigned char* buf[8];
};
__device__ __forceinline__ uint2* exchange_slot(unsigned char* buf, unsigned int parity, int rank) {
    constexpr size_t kHdr = 256, kPitch = 164944, kPeers = 8;
    return reinterpret_cast<uint2*>(buf + kHdr) + ((size_t)parity * kPeers + rank) * kPitch;
}
__device__ __forceinline__ void push_grad(const PushCtx& px, unsigned int tag, int off, float v) {
    // ONE naturally aligned 64-bit store: value in the low half, tag in the high half.  A b64 access is single-copy atomic
    // in the PTX memory model; a v2.u32 access is two scalar accesses whose order is unspecified (the receiver could see the
    // new tag beside a stale value).
    const unsigned long long word = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
#pragma unroll 1
    for (int p = 0; p < px.world; ++p) {
        uint2* dst = exchange_slot(px.buf[p], (tag - 1u) & 1u, px.rank) + off;
        asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
    }
}
// tag of the exchange in progress = epoch of the local buffer + 1 (stable during a step); 0 when there is no exchange
__device__ __forceinline__ unsigned int push_tag(const PushCtx& px) {
    return px.world > 1 ? *reinterpret_cast<const volatile unsigned int*>(px.buf[px.rank]) + 1u : 0u;
}

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in a fixed order (warp shuffles, then warp 0 adds the per-warp values in index
// order): deterministic for a given block size.  `scratch` holds >= 32 values.  Result valid in
// thread 0.  Ends with a barrier so scratch can be reused immediately.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    T r = T(0);
    if (wid == 0) {
        r = lane < nw ? scratch[lane] : T(0);
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// the four-tap lerp of the 256x256 resize with a fixed operation order (explicit FMAs), so that every kernel that evaluates it
// (resize256_kernel, the fused synthesis + resize pass) produces the same bits
__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, float lx, float ly) {
    const float hx = 1.f - lx, hy = 1.f - ly;
    const float top = __fmaf_rn(lx, v01, __fmul_rn(hx, v00)), bot = __fmaf_rn(lx, v11, __fmul_rn(hx, v10));
    return __fmaf_rn(ly, bot, __fmul_rn(hy, top));
}

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : kLeaky * v; }
__device__ __forceinline__ float leaky_grad(float act, float g) { return act > 0.f ? g : kLeaky * g; }

// reflect index of F.pad(mode='reflect'): -i -> i, (n-1)+i -> (n-1)-i   (valid for |overhang| <= n-1)
__device__ __forceinline__ int reflect(int i, int n) {
    i = i < 0 ? -i : i;
    return i > n - 1 ? 2 * (n - 1) - i : i;
}

}  // namespace dd
