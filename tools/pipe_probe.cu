#include <cuda_fp16.h>
#include <cstdio>
__global__ void __launch_bounds__(256) k16(__half2* out, __half2 a, __half2 b, int iters){
  __half2 x[8];
  for(int c=0;c<8;c++) x[c]=__hmul2(a,__floats2half2_rn((float)(threadIdx.x+c),(float)c));
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<8;c++){ __half2 d=__hsub2(x[c],a); x[c]=__hfma2(d,b,x[c]); }
  }
  __half2 s=x[0]; for(int c=1;c<8;c++) s=__hadd2(s,x[c]);
  out[blockIdx.x*256+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k32(float2* out, float a, float b, int iters){
  float2 x[8];
  for(int c=0;c<8;c++) x[c]=make_float2(a*(threadIdx.x+c),a*c);
  float2 aa=make_float2(a,a), bb=make_float2(b,b);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<8;c++){ float2 d=__fadd2_rn(x[c],aa); x[c]=__ffma2_rn(d,bb,x[c]); }
  }
  float2 s=x[0]; for(int c=1;c<8;c++) s=__fadd2_rn(s,x[c]);
  out[blockIdx.x*256+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k32s(float* out, float a, float b, int iters){
  float x[8];
  for(int c=0;c<8;c++) x[c]=a*(threadIdx.x+c);
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int c=0;c<8;c++){ float d=__fadd_rn(x[c],a); x[c]=__fmaf_rn(d,b,x[c]); }
  }
  float s=x[0]; for(int c=1;c<8;c++) s+=x[c];
  out[blockIdx.x*256+threadIdx.x]=s;
}
int main(){
  void* o; cudaMalloc(&o, 148*8*256*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=16384; float ms;
  for(int rep=0;rep<3;rep++){
    cudaEventRecord(e0); k16<<<148*8,256>>>((__half2*)o,__floats2half2_rn(1e-3f,1e-3f),__floats2half2_rn(1.001f,1.001f),iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("half2  : %.3f ms  %.2f T packed-instr-lanes/s (pairs of halves)\n", ms, 148.0*8*256*iters*16/ms/1e9);
    cudaEventRecord(e0); k32<<<148*8,256>>>((float2*)o,1e-3f,1.001f,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("float2 : %.3f ms  %.2f T packed-instr-lanes/s\n", ms, 148.0*8*256*iters*16/ms/1e9);
    cudaEventRecord(e0); k32s<<<148*8,256>>>((float*)o,1e-3f,1.001f,iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("float  : %.3f ms  %.2f T instr-lanes/s\n", ms, 148.0*8*256*iters*16/ms/1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
