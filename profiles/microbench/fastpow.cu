// accuracy of the MUFU-based pow used by the fused filter kernels: exp2f(g * __log2f(x)) vs double pow
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const float* x, float g, float* fast, float* acc, float* lfast, int n){
  int i = blockIdx.x*blockDim.x+threadIdx.x; if(i>=n) return;
  float l = __log2f(x[i]);
  fast[i] = exp2f(g*l);      // ex2.approx.ftz
  acc[i] = powf(x[i], g);
  lfast[i] = l;
}
int main(){
  const int n = 1<<22; float *hx = new float[n], *hf=new float[n], *ha=new float[n], *hl=new float[n];
  float *x,*f,*a,*l; cudaMalloc(&x,n*4); cudaMalloc(&f,n*4); cudaMalloc(&a,n*4); cudaMalloc(&l,n*4);
  // log-uniform in [1e-4, 4]
  for(int i=0;i<n;++i){ double u = (i+0.5)/n; hx[i] = (float)(1e-4*pow(4e4,u)); }
  cudaMemcpy(x,hx,n*4,cudaMemcpyHostToDevice);
  float gs[5]={1.f/3.f,0.7f,1.0f,1.7f,3.0f};
  for(float g: gs){
    k<<<n/256,256>>>(x,g,f,a,l,n); cudaMemcpy(hf,f,n*4,cudaMemcpyDeviceToHost); cudaMemcpy(ha,a,n*4,cudaMemcpyDeviceToHost); cudaMemcpy(hl,l,n*4,cudaMemcpyDeviceToHost);
    double mrel_f=0,mrel_a=0,mabs_f=0, mabs_l=0; 
    for(int i=0;i<n;++i){ double t = pow((double)hx[i],(double)g); double ef=fabs(hf[i]-t), ea=fabs(ha[i]-t);
      if(ef/t>mrel_f) mrel_f=ef/t; if(ea/t>mrel_a) mrel_a=ea/t; if(ef>mabs_f) mabs_f=ef; double el=fabs(hl[i]-log2((double)hx[i])); if(el>mabs_l) mabs_l=el; }
    printf("gamma %.3f: fast max rel %.3e  max abs %.3e | powf max rel %.3e | __log2f max abs err %.3e\n", g, mrel_f, mabs_f, mrel_a, mabs_l);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
