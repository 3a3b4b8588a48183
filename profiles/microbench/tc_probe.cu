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

template <int CIN, int COUT, int HIN>
static void probe_bwd(int B, int n_w) {
    constexpr int HO = HIN / 2;
    const size_t n_in = (size_t)B * CIN * HIN * HIN, n_out = (size_t)B * COUT * HO * HO;
    float *in, *dpre, *w, *din, *partial;
    cudaMalloc(&in, n_in * 4); cudaMalloc(&din, n_in * 4); cudaMalloc(&dpre, n_out * 4); cudaMalloc(&w, 32 * CIN * COUT * 4);
    cudaMalloc(&partial, (size_t)148 * 300 * 32 * 4);
    cudaMemset(in, 0, n_in * 4); cudaMemset(w, 0, 32 * CIN * COUT * 4); cudaMemset(dpre, 0, n_out * 4);
    auto kern = tc::conv_tc_bwd<CIN, COUT, HIN>;
    constexpr size_t smem = tc::conv_tc_bwd_smem<CIN, COUT, HIN>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int total = B * HO * HO;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        kern<<<148, 288, smem>>>(in, dpre, w, in, partial, din, n_w, total);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cy[256];
        cudaMemcpyFromSymbol(cy, tc::g_tc_cta_cycles, sizeof(cy));
        long long wmax = 0, wmin = 1ll << 60, dmax = 0, dmin = 1ll << 60;
        for (int i = 0; i < 148; ++i) {
            if (i < n_w) { wmax = cy[i] > wmax ? cy[i] : wmax; wmin = cy[i] < wmin ? cy[i] : wmin; }
            else { dmax = cy[i] > dmax ? cy[i] : dmax; dmin = cy[i] < dmin ? cy[i] : dmin; }
        }
        if (it == 2) {
            long long st[64];
            cudaMemcpyFromSymbol(st, tc::g_tc_stamp, sizeof(st));
            for (int g = 0; g < 8; ++g)
                printf("   dgrad stage %d: producer begin t=%lld split+st+arrive %lld | mma full-seen t=%lld issue %lld\n", g, st[2 * g] - st[63],
                       st[2 * g + 1] - st[2 * g], st[16 + 2 * g] - st[63], st[17 + 2 * g] - st[16 + 2 * g]);
            for (int k = 0; k < 2; ++k)
                printf("   epilogue %d: begin t=%lld done-wait %lld  ld+store %lld\n", k, st[32 + 3 * k] - st[63], st[33 + 3 * k] - st[32 + 3 * k],
                       st[34 + 3 * k] - st[33 + 3 * k]);
        }
        const int wt = (total + tc::WgradCfg<CIN>::PXT - 1) / tc::WgradCfg<CIN>::PXT, dt = (total + 127) / 128;
        printf("conv_tc_bwd<%d,%d,%d> B=%d n_w=%d: %s, %.1f us; wgrad CTA cycles [%lld, %lld] (%d tiles, %.1f per CTA); dgrad [%lld, %lld] (%d tiles, %.1f per CTA)\n",
               CIN, COUT, HIN, B, n_w, cudaGetErrorString(e), ms * 1e3f, wmin, wmax, wt, (double)wt / n_w, dmin, dmax, dt, (double)dt / (148 - n_w));
    }
    cudaFree(in); cudaFree(din); cudaFree(dpre); cudaFree(w); cudaFree(partial);
}

int main(int argc, char** argv) {
    if (argc > 1) {
        probe_bwd<16, 32, 128>(16, 60);
        probe_bwd<32, 32, 64>(16, 100);
        return 0;
    }
    probe<32, 32, 16>(16);
    probe<32, 32, 64>(16);
    probe<16, 32, 128>(16);
    return 0;
}
