// blur_tc_probe.cu -- in-kernel time stamps of the tensor-core filter kernels (dd_blur_tc.cuh), CTA 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDD_BTC_TIMING -I dedark_yolo_b200/csrc \
//        -o profiles/microbench/blur_tc_probe profiles/microbench/blur_tc_probe.cu && profiles/microbench/blur_tc_probe
// Prints, for the first blocks of CTA 0, how long the compute warps spend waiting for pass 1 / pass 2, splitting, in the
// epilogue, at the barrier and staging, and when the MMA warp issues what (cycles since kernel start).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dd_recovery_tc_fwd.cuh"
namespace dd {
void set_error(const char*, ...) {}
void count_launch(unsigned) {}
int check_launch(const char*) { return 0; }
}  // namespace dd
using namespace dd;

template <int R, bool X3, bool DBG>
static void probe(int B, int H, int W) {
    const size_t n = (size_t)B * 3 * H * W;
    float *x, *y, *feat;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&feat, B * 15 * 4);
    std::vector<float> hx(n), hf(B * 15);
    for (size_t i = 0; i < n; ++i) hx[i] = (float)rand() / RAND_MAX;
    for (auto& f : hf) f = 0.2f * ((float)rand() / RAND_MAX - 0.5f);
    cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(feat, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice);
    auto kern = btc::recovery_fwd_tc_kernel<R, X3, false, true, DBG>;
    constexpr size_t smem = btc::Lay<R, X3>::SMEM;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        kern<<<sc.G, btc::kThreadsTC, smem>>>(x, nullptr, nullptr, feat, y, B, H, W);
        cudaEventRecord(e1);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("recovery_fwd_tc_kernel<R=%d,X3=%d,DBG=%d> %dx3x%dx%d grid=%d smem=%zu: %s, %.1f us\n", R, (int)X3, (int)DBG, B, H, W, sc.G, smem,
               cudaGetErrorString(e), ms * 1e3f);
        if (it == 2) {
            long long st[512];
            cudaMemcpyFromSymbol(st, btc::g_btc_stamp, sizeof(st));
            const long long t0 = st[0];
            printf("  set-up %lld cycles, kernel %lld cycles\n", st[1] - t0, st[2] - t0);
            // iteration G of the compute warps: epilogue(G-1) [wait for pass 2, then the epilogue proper], barrier, loads of block G+2
            // (+ segment set-up) and wait for pass 1 of block G, split(G), chain + stores of block G+2
            printf("  compute warp 0:  blk   top    wait-p2  epilogue  barrier  load+wait-p1  split  store  | MMA warp: tile-wait-until  issue-p1  wait  issue-p2\n");
            for (int i = 0; i < 24; ++i) {
                const long long* c = st + 16 + 8 * i;
                const long long* m = st + 256 + 4 * i;
                printf("  %3d  t=%7lld  %7lld %8lld %7lld %12lld %6lld %6lld   | %7lld %8lld %6lld %8lld\n", i, c[0] - t0, i ? c[3] - c[0] : 0,
                       i ? c[4] - c[3] : c[4] - c[0], c[5] - c[4], c[1] - c[5], c[2] - c[1], c[6] - c[2], m[0] - t0, m[1] - m[0], m[2] - m[1], m[3] - m[2]);
            }
        }
    }
    cudaFree(x); cudaFree(y); cudaFree(feat);
}

int main() {
    probe<48, true, false>(16, 640, 640);
    probe<48, false, false>(16, 640, 640);
    probe<48, true, true>(16, 640, 640);
    return 0;
}
