// Microbenchmark: FP64 throughput of DFMA vs DMMA (mma.sync.m8n8k4.f64) on this GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dfma(double* out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-7;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 148 * 8 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int blocks = p.multiProcessorCount, threads = 32 * warps;
    float ms;
    k_dmma<8><<<blocks, threads>>>(d, 100);
    cudaEventRecord(e0); k_dmma<8><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tf_mma = 2.0 * 256 * 8 * (double)iters * blocks * warps / (ms * 1e-3) / 1e12;
    k_dfma<8><<<blocks, threads>>>(d, 100);
    cudaEventRecord(e0); k_dfma<8><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double tf_fma = 2.0 * 32 * 8 * (double)iters * blocks * warps / (ms * 1e-3) / 1e12;
    printf("%s SMs=%d warps/SM=%2d  DMMA %.2f TFLOP/s   DFMA %.2f TFLOP/s\n", p.name, blocks, warps, tf_mma, tf_fma);
  }
  // latency: single warp, dependent chain
  {
    float ms; int it2 = 200000;
    k_dmma<1><<<1, 32>>>(d, 100);
    cudaEventRecord(e0); k_dmma<1><<<1, 32>>>(d, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DMMA dependent-chain latency ~ %.1f ns per mma\n", ms * 1e6 / it2);
    k_dfma<1><<<1, 32>>>(d, 100);
    cudaEventRecord(e0); k_dfma<1><<<1, 32>>>(d, it2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA dependent-chain latency ~ %.1f ns per fma\n", ms * 1e6 / it2);
  }
  return 0;
}
