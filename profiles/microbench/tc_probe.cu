// tc_probe.cu -- in-kernel time stamps of the tensor-core forward conv (dd_conv_tc.cuh), CTA 0 / thread 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDD_TC_TIMING -I dedark_yolo_b200/csrc \
//        -o build/tc_probe profiles/microbench/tc_probe.cu && build/tc_probe
#include <cstdio>
#include <vector>
#include "dd_conv_tc.cuh"
namespace dd {
void set_error(const char*, ...) {}
void count_launch(unsigned) {}
int check_launch(const char*) { return 0; }
}  // namespace dd
using namespace dd;

template <int CIN, int COUT, int HIN>
static void probe(int B) {
    constexpr int HO = HIN / 2;
    const size_t n_in = (size_t)B * CIN * HIN * HIN, n_out = (size_t)B * COUT * HO * HO, n_w = 18 * CIN * COUT;
    float *in, *out, *w, *bias;
    cudaMalloc(&in, n_in * 4); cudaMalloc(&out, n_out * 4); cudaMalloc(&w, n_w * 4); cudaMalloc(&bias, COUT * 4);
    cudaMemset(in, 0, n_in * 4); cudaMemset(w, 0, n_w * 4); cudaMemset(bias, 0, COUT * 4);
    auto kern = tc::conv_tc_fwd<CIN, COUT, HIN>;
    constexpr size_t smem = tc::conv_tc_fwd_smem<CIN, COUT>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int total = B * HO * HO, ntl = (total + 127) / 128;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        kern<<<ntl < 148 ? ntl : 148, 288, smem>>>(in, w, bias, out, total);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long st[64];
        cudaMemcpyFromSymbol(st, tc::g_tc_stamp, sizeof(st));
        printf("conv_tc_fwd<%d,%d,%d> B=%d grid=%d: %s, %.1f us\n", CIN, COUT, HIN, B, (total + 127) / 128, cudaGetErrorString(e), ms * 1e3f);
        if (it == 2) {
            printf("  setup %lld\n", st[1] - st[0]);
            for (int g = 0; g < 8 && g < (ntl + 147) / 148 * (CIN / 8); ++g)
                printf("  stage %d: t=%lld  split+arrive %lld  gather-issue %lld\n", g, st[2 + 4 * g] - st[0], st[3 + 4 * g] - st[2 + 4 * g],
                       st[4 + 4 * g] - st[3 + 4 * g]);
            for (int g = 0; g < 8 && g < (ntl + 147) / 148 * (CIN / 8); ++g)
                printf("  mma warp stage %d: full seen t=%lld  issue %lld\n", g, st[44 + 2 * g] - st[0], st[45 + 2 * g] - st[44 + 2 * g]);
            printf("  last done-wait at t=%lld  epilogue %lld  total %lld cycles\n", st[40] - st[0], st[41] - st[40], st[42] - st[0]);
        }
    }
    cudaFree(in); cudaFree(out); cudaFree(w); cudaFree(bias);
}

int main() {
    probe<32, 32, 16>(16);
    probe<32, 32, 64>(16);
    probe<16, 32, 128>(16);
    return 0;
}
