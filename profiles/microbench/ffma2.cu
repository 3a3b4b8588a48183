#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void unpack(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c){ u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}

// variant A: 25-tap on 8 outputs with scalar FFMA-imm
__global__ void k_ffma(const float* in, float* out, int iters){
  float v[32]; for(int i=0;i<32;++i) v[i]=in[threadIdx.x*32+i];
  float acc[8]={0};
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int t=0;t<8;++t){
#pragma unroll
      for(int j=0;j<25;++j) acc[t]=fmaf(v[t+j], 0.01f*(j+1), acc[t]);
    }
#pragma unroll
    for(int i=0;i<8;++i) v[i]+=acc[i]*1e-9f;
  }
  for(int t=0;t<8;++t) out[threadIdx.x*8+t+blockIdx.x*blockDim.x*8]=acc[t];
}
// variant B: packed: two accumulators per output, adjacent input pairs, weight pairs as constants
__global__ void k_ffma2(const float* in, float* out, int iters){
  float v[32]; for(int i=0;i<32;++i) v[i]=in[threadIdx.x*32+i];
  float acc[8]={0};
  for(int it=0; it<iters; ++it){
    u64 P[16];
#pragma unroll
    for(int i=0;i<16;++i) P[i]=pack(v[2*i],v[2*i+1]);
#pragma unroll
    for(int t=0;t<8;++t){
      u64 a2 = pack(0.f,0.f);
      float s;
      if((t&1)==0){
#pragma unroll
        for(int j=0;j<24;j+=2) a2=ffma2(P[(t+j)/2], pack(0.01f*(j+1),0.01f*(j+2)), a2);
        s = v[t+24]*0.25f;
      } else {
#pragma unroll
        for(int j=1;j<25;j+=2) a2=ffma2(P[(t+j)/2], pack(0.01f*(j+1),0.01f*(j+2)), a2);
        s = v[t]*0.01f;
      }
      float x,y; unpack(a2,x,y);
      acc[t]+= x+y+s;
    }
#pragma unroll
    for(int i=0;i<8;++i) v[i]+=acc[i]*1e-9f;
  }
  for(int t=0;t<8;++t) out[threadIdx.x*8+t+blockIdx.x*blockDim.x*8]=acc[t];
}
// variant C: column pairs, same weight in both lanes (vertical-pass style): 4 pair-outputs x 25 taps
__global__ void k_ffma2v(const float* in, float* out, int iters){
  u64 P[32]; for(int i=0;i<32;++i) P[i]=pack(in[threadIdx.x*64+2*i], in[threadIdx.x*64+2*i+1]);
  u64 acc[8]; for(int i=0;i<8;++i) acc[i]=pack(0.f,0.f);
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int t=0;t<8;++t){
#pragma unroll
      for(int j=0;j<25;++j) acc[t]=ffma2(P[t+j], pack(0.01f*(j+1),0.01f*(j+1)), acc[t]);
    }
#pragma unroll
    for(int i=0;i<8;++i) P[i]=ffma2(acc[i], pack(1e-9f,1e-9f), P[i]);
  }
  for(int t=0;t<8;++t){ float x,y; unpack(acc[t],x,y); out[threadIdx.x*16+2*t+blockIdx.x*blockDim.x*16]=x; out[threadIdx.x*16+2*t+1+blockIdx.x*blockDim.x*16]=y; }
}
int main(){
  float *in,*out; cudaMalloc(&in, 1<<22); cudaMalloc(&out, 148*8*256*16*4*2); cudaMemset(in,0,1<<22);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=2000; float ms;
  for(int occ=1; occ<=4; occ*=2){
    int grid=148*occ*2, block=256;
    for(int rep=0;rep<2;++rep){
    cudaEventRecord(e0); k_ffma<<<grid,block>>>(in,out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    if(rep) printf("ffma   grid %d: %.3f ms  %.1f GFMA/s (200 fma/thread/iter)\n",grid,ms, 200.0*iters*grid*block/ms/1e6);
    cudaEventRecord(e0); k_ffma2<<<grid,block>>>(in,out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    if(rep) printf("ffma2h grid %d: %.3f ms  %.1f GFMA/s (200 fma-equiv)\n",grid,ms, 200.0*iters*grid*block/ms/1e6);
    cudaEventRecord(e0); k_ffma2v<<<grid,block>>>(in,out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    if(rep) printf("ffma2v grid %d: %.3f ms  %.1f GFMA/s (400 fma-equiv)\n",grid,ms, 400.0*iters*grid*block/ms/1e6);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
