// FP64 FMA peak probe: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double b = 1.0000001, c = 0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    int blocks = 148 * (2048 / threads);
    int iters = 200000;
    k<<<blocks, threads>>>(d, 1000); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * (double)iters * blocks * threads;
    printf("threads=%d blocks=%d: %.2f ms, %.2f TFLOP/s FP64\n", threads, blocks, ms, fl / ms / 1e9);
  }
  return 0;
}
