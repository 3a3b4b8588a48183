// Can a CTA that needs ~200 KB of shared memory start on an SM that is running a no-shared-memory kernel of another stream?
// A: 1 CTA/SM spinning for ~200 us (no smem).  B: 148 CTAs x 200 KB dynamic smem, trivial.  B is launched right after A on a
// second stream; we time B's completion.  Variants: A with the default carve-out, A with PreferredSharedMemoryCarveout = max,
// A with a token 1 KB dynamic smem + carve-out max.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(long long cycles, int* sink) {
    extern __shared__ int dummy[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 9999) sink[0] = dummy[0];
}
__global__ void bigsmem(int* out) {
    extern __shared__ int s[];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[31];
}
int main() {
    int* out; cudaMalloc(&out, 4096);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaFuncSetAttribute(bigsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, eA, eB; cudaEventCreate(&e0); cudaEventCreate(&eA); cudaEventCreate(&eB);
    for (int variant = 0; variant < 4; ++variant) {
        int carve = variant == 0 ? -1 : 100;
        size_t dyn = variant == 2 ? 1024 : (variant == 3 ? 16 * 1024 : 0);
        cudaFuncSetAttribute(spin, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        for (int rep = 0; rep < 3; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0, 0);
            spin<<<148, 256, dyn, s1>>>(400000, nullptr);  // ~200 us
            cudaEventRecord(eA, s1);
            bigsmem<<<148, 256, 200 * 1024, s2>>>(out);
            cudaEventRecord(eB, s2);
            cudaDeviceSynchronize();
            float a, b; cudaEventElapsedTime(&a, e0, eA); cudaEventElapsedTime(&b, e0, eB);
            if (rep == 2) printf("variant %d (carveout %d, dyn %zu): spin done at %.1f us, big-smem kernel done at %.1f us  [%s]\n", variant, carve, dyn, a * 1e3, b * 1e3, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
