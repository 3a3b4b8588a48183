// blur_tc_probe_one.cu -- one launch of the tensor-core forward filter kernel on 16x3x640x640, for `ncu --set full`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I dedark_yolo_b200/csrc \
//        -o profiles/microbench/blur_tc_one_probe profiles/microbench/blur_tc_probe_one.cu
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
int main() {
    const int B = 16, H = 640, W = 640;
    const size_t n = (size_t)B * 3 * H * W;
    float *x, *y, *feat;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMalloc(&feat, B * 15 * 4);
    std::vector<float> hx(n), hf(B * 15);
    for (size_t i = 0; i < n; ++i) hx[i] = (float)rand() / RAND_MAX;
    for (auto& f : hf) f = 0.2f * ((float)rand() / RAND_MAX - 0.5f);
    cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(feat, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice);
    auto kern = btc::recovery_fwd_tc_kernel<48, true, false, true, false>;
    constexpr size_t smem = btc::Lay<48, true>::SMEM;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const Sched sc = make_sched(B, H, W, btc::kSchedCtasTC);
    for (int it = 0; it < 3; ++it) kern<<<sc.G, btc::kThreadsTC, smem>>>(x, nullptr, nullptr, feat, y, B, H, W);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
