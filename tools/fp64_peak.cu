// FP64 FMA peak probe: register operands vs constant-bank / uniform-register operands (as the tile kernel uses them)
#include <cstdio>
#include <cuda_runtime.h>
struct Coef { double v[32]; };
__global__ void k_reg(double* out, int iters) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double b = 1.0000001, c = 0.5;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_const(double* out, int iters, const __grid_constant__ Coef C) {
  double a[8], x[8];
  for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x + j; x[j] = 1.0 + 1e-9 * (threadIdx.x + j); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = fma(x[(j + k) & 7], C.v[(k * 8 + j) & 31], a[j]);
  }
  double s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* d; cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  Coef C; for (int i = 0; i < 32; ++i) C.v[i] = 1e-7 * (i + 1);
  for (int threads : {128, 256, 1024}) {
    for (int occ : {3, 8}) {
      int blocks = 148 * (threads == 1024 ? 2 : occ);
      int iters = 100000;
      float ms;
      k_reg<<<blocks, threads>>>(d, 1000); cudaDeviceSynchronize();
      cudaEventRecord(e0); k_reg<<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("reg   operands threads=%4d blocks/SM=%d: %.2f TFLOP/s\n", threads, blocks / 148, 2.0 * 8 * iters * blocks * threads / ms / 1e9);
      k_const<<<blocks, threads>>>(d, 1000, C); cudaDeviceSynchronize();
      cudaEventRecord(e0); k_const<<<blocks, threads>>>(d, iters / 4, C); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("const operands threads=%4d blocks/SM=%d: %.2f TFLOP/s\n", threads, blocks / 148, 2.0 * 32 * (iters / 4) * (double)blocks * threads / ms / 1e9);
    }
  }
  return 0;
}
