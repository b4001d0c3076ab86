// FP64 FMA peak of the device (the denominator of the assembly roofline):
// every SM full of warps, 8 independent DFMA chains per thread, timed with CUDA events.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024)
dfma_kernel(double *out, double a, double b, int iters)
{
  double c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    c[k] = a + k + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          c[k] = fma(c[k], b, a);
    }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    s += c[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int
main()
{
  int sm = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sm * 2, threads = 1024, iters = 4096;
  double   *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
    {
      cudaEventRecord(e0);
      dfma_kernel<<<blocks, threads>>>(out, 1.0000001, 0.9999999, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best)
        best = ms;
    }
  const double flop = 2.0 * 32 * iters * (double)blocks * threads;
  printf("{\"fp64_fma_tflops\": %.3f, \"sms\": %d, \"ms\": %.3f, \"how\": \"8 independent DFMA chains per thread, %d x %d threads, best of 4\"}\n",
         flop / (best * 1e-3) / 1e12, sm, best, blocks, threads);
  return 0;
}
