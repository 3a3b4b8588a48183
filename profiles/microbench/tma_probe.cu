// tma_probe.cu -- (1) which error does __trap() raise; (2) a 3-D TMA box with odd height / negative origin into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dedark_yolo_b200/csrc -o build/tma_probe profiles/microbench/tma_probe.cu
#include <cstdio>
#include <vector>
#include "dd_async.cuh"
namespace dd {
void set_error(const char*, ...) {}
void count_launch(unsigned) {}
int check_launch(const char*) { return 0; }
}  // namespace dd
using namespace dd;

__global__ void trap_kernel() { __trap(); }

template <int BW, int BH, int BD>
__global__ void tma_kernel(const __grid_constant__ CUtensorMap map, int x, int y, int z, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, BW * BH * BD * 4);
        tma_load_3d(smem, &map, x, y, z, &bar);
    }
    mbar_wait(&bar, 0);
    const float* s = reinterpret_cast<const float*>(smem);
    for (int i = threadIdx.x; i < BW * BH * BD; i += blockDim.x) out[i] = s[i];
}

template <int BW, int BH, int BD>
static void run(int x, int y, int z) {
    const int P = 6, H = 256, W = 256;
    std::vector<float> h((size_t)P * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * BD * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const bool ok = make_tensor_map_3d(&map, d, P, H, W, BW, BH, BD);
    auto kern = tma_kernel<BW, BH, BD>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BW * BH * BD * 4 + 128);
    kern<<<1, 256, BW * BH * BD * 4 + 128>>>(map, x, y, z, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> r(BW * BH * BD);
    cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int dz = 0; dz < BD; ++dz)
        for (int dy = 0; dy < BH; ++dy)
            for (int dx = 0; dx < BW; ++dx) {
                const int gx = x + dx, gy = y + dy, gz = z + dz;
                const float want = (gx < 0 || gx >= W || gy < 0 || gy >= H || gz < 0 || gz >= P) ? 0.f : h[((size_t)gz * H + gy) * W + gx];
                bad += r[(dz * BH + dy) * BW + dx] != want;
            }
    printf("box %dx%dx%d at (%d,%d,%d): encode %s, %s, %d mismatches\n", BW, BH, BD, x, y, z, ok ? "ok" : "FAILED", cudaGetErrorString(e), bad);
    cudaFree(d); cudaFree(o);
}

int main(int argc, char** argv) {
    if (argc > 1) {
        trap_kernel<<<1, 32>>>();
        printf("__trap(): %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    run<156, 16, 1>(-12, -12, 0);
    run<72, 33, 1>(-4, -1, 2);
    run<72, 33, 1>(188, 31, 2);
    run<72, 33, 3>(-4, -1, 3);
    run<72, 17, 3>(60, 15, 0);
    run<32, 8, 16>(32, 8, 0);
    run<68, 33, 1>(-1, -1, 2);  // innermost origin not 16-byte aligned: illegal instruction (poisons the context: keep last)
    return 0;
}
