// Microbenchmark: latency of one producer->consumer hop through L2 on B200
// (st.relaxed.gpu by one warp, ld.relaxed.gpu polling by another on a different SM).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ldv(const unsigned long long *p)
{ unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stv(unsigned long long *p, unsigned long long v)
{ asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// chain element i is handled by block (i % gridDim.x); it waits for x[i-1] != 0 and writes x[i].
__global__ void chain(unsigned long long *x, int n, int mode)
{
  for (int i = blockIdx.x + 1; i <= n; i += gridDim.x)
    {
      if (mode == 0) { if (threadIdx.x == 0) { while (ldv(x + i - 1) == 0) {} stv(x + i, i + 1); } }
      else {  // all 32 lanes poll (same address), shuffle-reduce 5 rounds, lane 0 stores
        unsigned long long v; while ((v = ldv(x + i - 1)) == 0) {}
        double s = (double)v;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) stv(x + i, (unsigned long long)s | 1ull);
      }
    }
}
int main()
{
  const int n = 20000;
  unsigned long long *x; cudaMalloc(&x, (n + 1) * 8);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {2, 8, 32, 148, 592})
      {
        cudaMemset(x, 0, (n + 1) * 8);
        unsigned long long one = 1; cudaMemcpy(x, &one, 8, cudaMemcpyHostToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); chain<<<grid, 32>>>(x, n, mode); cudaEventRecord(b); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("mode %d grid %d: %.3f us per hop\n", mode, grid, ms * 1e3 / n);
      }
  return 0;
}
