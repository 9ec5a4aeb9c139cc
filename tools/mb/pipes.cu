// Pipe microbenchmark: IMAD.WIDE alone, DFMA alone, both interleaved in one instruction stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; prints G warp-lane ops/s
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint64_t* out, double* dout, int iters) {
  uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7;
  uint64_t x0 = a, x1 = b, x2 = a ^ b, x3 = a + b;
  double d0 = 1.0 + threadIdx.x * 1e-9, d1 = 1.5, d2 = 0.25 + blockIdx.x * 1e-9, d3 = 3.0;
  const double m = 1.0000000001, c = 1e-30;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0 || MODE == 2) {
        x0 = (uint64_t)(uint32_t)x0 * a + x0;
        x1 = (uint64_t)(uint32_t)x1 * b + x1;
        x2 = (uint64_t)(uint32_t)x2 * a + x2;
        x3 = (uint64_t)(uint32_t)x3 * b + x3;
      }
      if (MODE == 1 || MODE == 2) {
        d0 = fma(d0, m, c);
        d1 = fma(d1, m, c);
        d2 = fma(d2, m, c);
        d3 = fma(d3, m, c);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3;
  dout[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
}
int main() {
  int sm = 148;
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  sm = p.multiProcessorCount;
  const int blocks = sm * 8, threads = 256, iters = 2000;
  uint64_t* o;
  double* d;
  cudaMalloc(&o, blocks * threads * 8);
  cudaMalloc(&d, blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<blocks, threads>>>(o, d, iters);
      if (mode == 1) k<1><<<blocks, threads>>>(o, d, iters);
      if (mode == 2) k<2><<<blocks, threads>>>(o, d, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double ops = (double)blocks * threads * iters * 16 * 4;
    printf("mode %d (%s): %.3f ms, %.1f G lane-ops/s per kind\n", mode,
           mode == 0 ? "IMAD.WIDE" : (mode == 1 ? "DFMA" : "IMAD.WIDE + DFMA interleaved"), ms, ops / ms / 1e6);
  }
  return 0;
}
