// Microbenchmarks that fix the roofline denominators for the FP64 path on B200:
//   * raw DMMA (mma.sync f64) issue rate per shape / warps / ILP
//   * raw DFMA rate
//   * cuBLAS DGEMM / ZGEMM ceilings at 8192^3 and at the H_eff GEMM shapes
//   * device copy bandwidth
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuComplex.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP, int SHAPE>
__global__ void __launch_bounds__(1024) dmma_rate(double* out, int iters){
  double a[8], b[4], c[ILP][4];
  for(int i=0;i<8;i++) a[i]=1.0+threadIdx.x*1e-9+i;
  for(int i=0;i<4;i++) b[i]=1e-9*(threadIdx.x+i);
  for(int j=0;j<ILP;j++) for(int i=0;i<4;i++) c[j][i]=0.0;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int j=0;j<ILP;j++){
      if(SHAPE==16)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};\n"
        : "+d"(c[j][0]),"+d"(c[j][1]),"+d"(c[j][2]),"+d"(c[j][3]) : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
      else if(SHAPE==8)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};\n"
        : "+d"(c[j][0]),"+d"(c[j][1]),"+d"(c[j][2]),"+d"(c[j][3]) : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
      else if(SHAPE==4)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};\n"
        : "+d"(c[j][0]),"+d"(c[j][1]),"+d"(c[j][2]),"+d"(c[j][3]) : "d"(a[0]),"d"(a[1]),"d"(b[0]));
      else
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
        : "+d"(c[j][0]),"+d"(c[j][1]) : "d"(a[0]),"d"(b[0]));
    }
  }
  double s=0; for(int j=0;j<ILP;j++) for(int i=0;i<4;i++) s+=c[j][i];
  if(s==123.456) out[0]=s;
}

template<int ILP>
__global__ void __launch_bounds__(1024) dfma_rate(double* out, int iters){
  double c[ILP]; double a=1.0+1e-9*threadIdx.x, b=1e-9*threadIdx.x;
  for(int j=0;j<ILP;j++) c[j]=j;
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int j=0;j<ILP;j++) c[j]=fma(a,c[j],b);
  }
  double s=0; for(int j=0;j<ILP;j++) s+=c[j];
  if(s==123.456) out[0]=s;
}

template<typename F> float time_ms(F f, int reps=3){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
  return best;
}

template<int ILP,int SHAPE> void run_dmma(double* d, int nsm, int warps){
  int iters=4096; int blocks=nsm; 
  float ms=time_ms([&]{ dmma_rate<ILP,SHAPE><<<blocks,warps*32>>>(d,iters); });
  double fl_per = SHAPE==16? 16.*8*16*2 : SHAPE==8? 16.*8*8*2 : SHAPE==4? 16.*8*4*2 : 8.*8*4*2;
  double fl=fl_per*ILP*(double)iters*warps*blocks;
  printf("{\"bench\":\"dmma\",\"shape\":%d,\"ilp\":%d,\"warps_per_sm\":%d,\"tflops\":%.3f,\"ms\":%.4f}\n",SHAPE,ILP,warps,fl/ms*1e-9,ms);
}
template<int ILP> void run_dfma(double* d,int nsm,int warps){
  int iters=8192; float ms=time_ms([&]{ dfma_rate<ILP><<<nsm,warps*32>>>(d,iters); });
  double fl=2.0*ILP*(double)iters*warps*32*nsm;
  printf("{\"bench\":\"dfma\",\"ilp\":%d,\"warps_per_sm\":%d,\"tflops\":%.3f}\n",ILP,warps,fl/ms*1e-9);
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int nsm=p.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"cc\":\"%d.%d\",\"smem_optin\":%zu,\"l2\":%d}\n",p.name,nsm,p.major,p.minor,p.sharedMemPerBlockOptin,p.l2CacheSize);
  double* d; CK(cudaMalloc(&d,1<<20));
  int ws[4]={4,8,16,32};
  for(int wi=0;wi<4;wi++){ int w=ws[wi];
    run_dmma<1,884>(d,nsm,w); run_dmma<4,884>(d,nsm,w); run_dmma<8,884>(d,nsm,w);
    run_dmma<1,4>(d,nsm,w); run_dmma<4,4>(d,nsm,w);
    run_dmma<1,8>(d,nsm,w); run_dmma<4,8>(d,nsm,w); run_dmma<8,8>(d,nsm,w);
    run_dmma<1,16>(d,nsm,w); run_dmma<4,16>(d,nsm,w); run_dmma<8,16>(d,nsm,w);
    run_dfma<4>(d,nsm,w); run_dfma<8>(d,nsm,w);
  }
  // copy bandwidth
  { size_t n=(size_t)1<<30; double *a,*b; CK(cudaMalloc(&a,n)); CK(cudaMalloc(&b,n)); CK(cudaMemset(a,1,n));
    float ms=time_ms([&]{ cudaMemcpyAsync(b,a,n,cudaMemcpyDeviceToDevice); },5);
    printf("{\"bench\":\"d2d_copy\",\"gbs\":%.1f}\n",2.0*n/ms*1e-6); cudaFree(a); cudaFree(b); }
  // pinned H2D / D2H
  { size_t n=(size_t)512<<20; void *h,*dv; CK(cudaMallocHost(&h,n)); CK(cudaMalloc(&dv,n));
    float ms=time_ms([&]{ cudaMemcpyAsync(dv,h,n,cudaMemcpyHostToDevice); },3);
    float ms2=time_ms([&]{ cudaMemcpyAsync(h,dv,n,cudaMemcpyDeviceToHost); },3);
    printf("{\"bench\":\"pcie\",\"h2d_gbs\":%.1f,\"d2h_gbs\":%.1f}\n",n/ms*1e-6,n/ms2*1e-6); cudaFreeHost(h); cudaFree(dv); }
  cublasHandle_t h; cublasCreate(&h);
  struct S{int m,n,k; cublasOperation_t ta,tb; const char* name;};
  S shapes[]={{8192,8192,8192,CUBLAS_OP_N,CUBLAS_OP_N,"sq8192_NN"},{8192,8192,8192,CUBLAS_OP_T,CUBLAS_OP_N,"sq8192_TN"},
              {20480,16384,4096,CUBLAS_OP_T,CUBLAS_OP_N,"K1_chi4096_TN"},{16384,4096,20480,CUBLAS_OP_N,CUBLAS_OP_N,"K3_chi4096_NN"},
              {10240,8192,2048,CUBLAS_OP_T,CUBLAS_OP_N,"K1_chi2048_TN"},{8192,2048,10240,CUBLAS_OP_N,CUBLAS_OP_N,"K3_chi2048_NN"},
              {5120,4096,1024,CUBLAS_OP_T,CUBLAS_OP_N,"K1_chi1024_TN"},{4096,1024,5120,CUBLAS_OP_N,CUBLAS_OP_N,"K3_chi1024_NN"},
              {2560,2048,512,CUBLAS_OP_T,CUBLAS_OP_N,"K1_chi512_TN"},{2048,512,2560,CUBLAS_OP_N,CUBLAS_OP_N,"K3_chi512_NN"}};
  for(auto&s:shapes){
    size_t na=(size_t)s.m*s.k, nb=(size_t)s.k*s.n, nc=(size_t)s.m*s.n; double *A,*B,*C;
    CK(cudaMalloc(&A,na*8)); CK(cudaMalloc(&B,nb*8)); CK(cudaMalloc(&C,nc*8));
    CK(cudaMemset(A,0,na*8)); CK(cudaMemset(B,0,nb*8));
    double al=1,be=0; int lda=s.ta==CUBLAS_OP_N?s.m:s.k, ldb=s.tb==CUBLAS_OP_N?s.k:s.n;
    float ms=time_ms([&]{ cublasDgemm(h,s.ta,s.tb,s.m,s.n,s.k,&al,A,lda,B,ldb,&be,C,s.m); },5);
    printf("{\"bench\":\"cublas_dgemm\",\"name\":\"%s\",\"m\":%d,\"n\":%d,\"k\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",s.name,s.m,s.n,s.k,ms,2.0*s.m*s.n*s.k/ms*1e-9);
    cudaFree(A);cudaFree(B);cudaFree(C);
  }
  { // sustained 4 s
    int n=8192; double *A,*B,*C; CK(cudaMalloc(&A,(size_t)n*n*8)); CK(cudaMalloc(&B,(size_t)n*n*8)); CK(cudaMalloc(&C,(size_t)n*n*8));
    CK(cudaMemset(A,0,(size_t)n*n*8)); CK(cudaMemset(B,0,(size_t)n*n*8)); double al=1,be=0;
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); int reps=120; cudaEventRecord(e0);
    for(int i=0;i<reps;i++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&al,A,n,B,n,&be,C,n);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1);
    printf("{\"bench\":\"cublas_dgemm_sustained\",\"n\":8192,\"reps\":%d,\"total_ms\":%.1f,\"tflops\":%.2f}\n",reps,ms,2.0*n*n*(double)n*reps/ms*1e-9);
    cudaFree(A);cudaFree(B);cudaFree(C); }
  { int m=5120,n=4096,k=1024; cuDoubleComplex *A,*B,*C; CK(cudaMalloc(&A,(size_t)m*k*16)); CK(cudaMalloc(&B,(size_t)k*n*16)); CK(cudaMalloc(&C,(size_t)m*n*16));
    CK(cudaMemset(A,0,(size_t)m*k*16)); CK(cudaMemset(B,0,(size_t)k*n*16)); cuDoubleComplex al={1,0},be={0,0};
    float ms=time_ms([&]{ cublasZgemm(h,CUBLAS_OP_T,CUBLAS_OP_N,m,n,k,&al,A,k,B,k,&be,C,m); },5);
    printf("{\"bench\":\"cublas_zgemm\",\"name\":\"K1_chi1024_TN\",\"ms\":%.3f,\"real_tflops\":%.2f}\n",ms,8.0*m*n*k/ms*1e-9); }
  return 0;
}
