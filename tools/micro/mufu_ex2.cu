// Microbenchmark: throughput of ex2.approx.ftz.f32 (MUFU.EX2) per SM sub-partition, and of the FFMA + EX2 + FADD + F2FP mix of the
// attention transform.  One CTA per SM, W warps per sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared (a statically linked cudart must not ship to the GPU box)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MIX>
__global__ void k(int iters, long long* out, float* sink, float seed) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  float rs = 0.f;
  uint32_t pk = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = MIX ? fmaf(x[i], 1.0001f, -0.3f) : x[i];
      float e;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
      if (MIX) {
        rs += e;
        if (i & 1) {
          __nv_bfloat162 t = __floats2bfloat162_rn(x[i - 1], e);
          pk ^= *reinterpret_cast<uint32_t*>(&t);
        }
      }
      x[i] = e * 0.5f;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  float s = rs + __uint_as_float(pk);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int mix = 0; mix < 2; ++mix)
    for (int warps : {4, 8, 16}) {
      if (mix) k<1><<<148, warps * 32>>>(iters, out, sink, 0.1f); else k<0><<<148, warps * 32>>>(iters, out, sink, 0.1f);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      const double n = (double)iters * 16 * (warps / 4);   // ex2 warp-instructions per sub-partition
      printf("%s, %2d warps/SM (%d per sub-partition): %.2f cycles per ex2 warp-instruction per sub-partition -> %.1f exp/clk/SM\n",
             mix ? "FFMA+EX2+FADD+F2FP mix" : "EX2 only", warps, warps / 4, c / n, 32.0 * 4 * n / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
