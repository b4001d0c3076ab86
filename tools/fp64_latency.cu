// Dependent-issue latencies that bound the triangular-solve recurrence (one warp):
// DFMA, DADD, DMUL chains, 64-bit shuffle, shared-memory load, and FP64 throughput per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double *out, long long *cyc, double a, double b)
{
  __shared__ double sh[64];
  const int lane = threadIdx.x;
  sh[lane] = a + lane; sh[lane + 32] = b;
  __syncthreads();
  double x = a + lane;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        x = fma(x, b, a);
    }
  long long t1 = clock64();
  double y = x;
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        y = y + b;
    }
  long long t2 = clock64();
  double z = y;
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        z = __shfl_xor_sync(0xffffffffu, z, 1 + (k & 15));
    }
  long long t3 = clock64();
  int idx = lane;
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        idx = (int)sh[idx & 63] & 31;
    }
  long long t4 = clock64();
  // 8 independent DFMA chains: throughput of one warp
  double c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k] = a + k;
#pragma unroll 1
  for (int i = 0; i < 64; ++i)
    {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          c[k] = fma(c[k], b, a);
    }
  long long t5 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z + idx + s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    {
      cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
    }
}

int main()
{
  double *out; long long *cyc, h[5];
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  for (int r = 0; r < 2; ++r)
    lat_kernel<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("one warp, cycles per dependent op: DFMA %.1f  DADD %.1f  SHFL64 %.1f  LDS.64->addr %.1f | 8 independent DFMA chains: %.1f cycles per warp-DFMA\n",
         h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0);
  return 0;
}
