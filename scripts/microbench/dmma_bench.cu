#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double *out, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1e-9 * threadIdx.x, b = 0.5;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps_per_sm, int sms) {
    double *out; cudaMalloc(&out, 1 << 24);
    int iters = 20000;
    int threads = 32 * warps_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC><<<sms, threads>>>(out, 10);
    cudaEventRecord(e0);
    k<NACC><<<sms, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)sms * warps_per_sm * iters * NACC;
    printf("NACC=%d warps/SM=%d: %.2f TFLOP/s, %.2f cycles/DMMA/SM at 1.965GHz (per-warp issue interval %.1f cyc)\n", NACC, warps_per_sm, n * 512 / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (iters * NACC * warps_per_sm), ms * 1e-3 * 1.965e9 / (iters * NACC));
    cudaFree(out);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1>(1, sms); run<2>(1, sms); run<4>(1, sms); run<8>(1, sms); run<11>(1, sms);
    run<8>(4, sms); run<8>(8, sms); run<8>(16, sms); run<11>(22, sms); run<4>(32, sms);
    return 0;
}
