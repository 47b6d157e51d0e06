// micro-benchmark: issue throughput of packed fp32 (FFMA2/FADD2/FMUL2) vs scalar FFMA on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template<int MODE> __global__ void __launch_bounds__(1024) k(float* out, int iters, float s){
  float a[16]; u64 p[8];
  for(int i=0;i<16;i++) a[i]=threadIdx.x*0.001f+i;
  for(int i=0;i<8;i++){ float2 t=make_float2(a[2*i],a[2*i+1]); p[i]=*reinterpret_cast<u64*>(&t);} 
  float2 sc2=make_float2(s,s); u64 sc=*reinterpret_cast<u64*>(&sc2);
  for(int it=0; it<iters; ++it){
    if(MODE==0){
#pragma unroll
      for(int i=0;i<16;i++) a[i]=__fmaf_rn(a[i],s,s);
    } else if(MODE==1){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],sc,sc);
    } else if(MODE==2){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=add2(p[i],sc);
    } else if(MODE==3){
#pragma unroll
      for(int i=0;i<8;i++) p[i]=mul2(p[i],sc);
    } else if(MODE==4){  // scalar FADD
#pragma unroll
      for(int i=0;i<16;i++) a[i]=__fadd_rn(a[i],s);
    } else if(MODE==5){  // mix: 8 FFMA2 + 8 IADD-like (int) to see co-issue
#pragma unroll
      for(int i=0;i<8;i++) p[i]=fma2(p[i],sc,sc);
#pragma unroll
      for(int i=0;i<8;i++) a[i]=__int_as_float(__float_as_int(a[i])+it);
    } else if(MODE==6){  // mix: 16 FFMA + 8 int
#pragma unroll
      for(int i=0;i<16;i++) a[i+0]=__fmaf_rn(a[i],s,s);
    }
  }
  float r=0; for(int i=0;i<16;i++) r+=a[i]; for(int i=0;i<8;i++){ float2 t=*reinterpret_cast<float2*>(&p[i]); r+=t.x+t.y; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> void run(const char* name, int inst_per_iter, float* d){
  int iters=8192; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148,1024>>>(d,iters,1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148,1024>>>(d,iters,1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double winst = 148.0*32*iters*inst_per_iter; // warp-insts
  printf("%-28s %8.3f ms  %7.2f warp-inst/ns total -> %.2f inst/clk/SM @1.965GHz\n", name, ms, winst/ms/1e6, winst/ms/1e6/148/1.965);
}
int main(){ float* d; cudaMalloc(&d,148*1024*4);
  run<0>("FFMA x16",16,d); run<1>("FFMA2 x8",8,d); run<2>("FADD2 x8",8,d); run<3>("FMUL2 x8",8,d); run<4>("FADD x16",16,d); run<5>("FFMA2 x8 + IADD x8",16,d);
  return 0; }
