#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dfma(double *out, long long *cyc) { double a = threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) a = fma(a, b, c); long long t1 = clock64(); out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_dmul_dadd(double *out, long long *cyc) { double a = threadIdx.x * 1e-9 + 1, b = 1.0000001; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) { a = a * b; a = a + b; } long long t1 = clock64(); out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 2; }
__global__ void k_lds(double *out, long long *cyc) { __shared__ int idx[1024]; for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 37 + 11) & 1023; __syncthreads(); int j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = idx[j]; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_ldgen(double *out, long long *cyc, int *gidx) { __shared__ int idx[1024]; for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 37 + 11) & 1023; __syncthreads(); 
  int *p = out ? idx : gidx; int j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = p[j]; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_ldg(double *out, long long *cyc, int *gidx) { int j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = gidx[j]; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_shfl(double *out, long long *cyc) { int j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = __shfl_sync(0xffffffffu, j, (j + 1) & 31); long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_shfl64(double *out, long long *cyc) { double j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = __shfl_sync(0xffffffffu, j, (i + 1) & 31) + 1.0; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_redux(double *out, long long *cyc) { unsigned j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = __reduce_max_sync(0xffffffffu, j ^ i) + threadIdx.x; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_ballot(double *out, long long *cyc) { unsigned j = threadIdx.x; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) j = __ballot_sync(0xffffffffu, (j ^ i) & 1) + threadIdx.x; long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_bar(double *out, long long *cyc) { long long t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads(); long long t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_iadd(double *out, long long *cyc) { int j = threadIdx.x, k = 3; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) { j = j * k + i; } long long t1 = clock64(); out[threadIdx.x] = j; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_fsel(double *out, long long *cyc) { double a = threadIdx.x, b = 2.0; long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < N; ++i) { a = (a > b) ? a - 1.0 : a + 2.0; } long long t1 = clock64(); out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
__global__ void k_rcp(double *out, long long *cyc) { double a = threadIdx.x + 1.5; long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < N; ++i) { a = 1.0 / a + 0.5; } long long t1 = clock64(); out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
int main() { double *out; long long *cyc; int *gidx; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64); cudaMalloc(&gidx, 4096);
  int h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023; cudaMemcpy(gidx, h, 4096, cudaMemcpyHostToDevice);
  long long c;
#define RUN(name, threads, ...) name<<<1, threads>>>(__VA_ARGS__); name<<<1, threads>>>(__VA_ARGS__); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-12s threads %4d: %.1f cycles/op\n", #name, threads, (double)c / N);
  RUN(k_dfma, 32, out, cyc) RUN(k_dmul_dadd, 32, out, cyc) RUN(k_fsel, 32, out, cyc) RUN(k_iadd, 32, out, cyc) RUN(k_lds, 32, out, cyc) RUN(k_ldgen, 32, out, cyc, gidx) RUN(k_ldg, 32, out, cyc, gidx)
  RUN(k_shfl, 32, out, cyc) RUN(k_shfl64, 32, out, cyc) RUN(k_redux, 32, out, cyc) RUN(k_ballot, 32, out, cyc) RUN(k_rcp, 32, out, cyc)
  RUN(k_bar, 32, out, cyc) RUN(k_bar, 352, out, cyc) RUN(k_bar, 1024, out, cyc) RUN(k_dfma, 352, out, cyc) RUN(k_dfma, 1024, out, cyc)
  return 0; }
