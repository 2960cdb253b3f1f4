// Microbenchmark: tcgen05.ld throughput per SM (how many bytes per clock TMEM -> registers), 1..8 warps loading concurrently.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int iters, int nwarps, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + (uint32_t)((i & 7) * 32))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int nw : {1, 2, 4, 8}) {
    k<<<148, 256>>>(iters, nw, out, sink);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)nw * iters * 32 * 32 * 4;
    printf("warps %d: %lld cycles for %d x32 loads per warp -> %.1f B/clk/SM (%.1f cycles per 4 KB warp load)\n", nw, c, iters, bytes / c,
           (double)c / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
