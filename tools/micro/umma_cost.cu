// Cost of one tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of where A comes from (shared memory / tensor memory),
// issued back to back by one thread: what a fused-attention tile pays per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cudart shared -I prompt_tts_b200/csrc -o tools/micro/umma_cost tools/micro/umma_cost.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"
using namespace tc;

// mode 0: A and B from shared memory, K-major SWIZZLE_128B (64-element rows);  mode 1: A from TMEM, B MN-major from shared memory
template <int N, int MODE>
__global__ void __launch_bounds__(128) k(long long* out, int reps, int per_commit) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 128 * 128, bar = base + 128 * 128 + 256 * 128, slot = bar + 16;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(raw + (slot - smem_u32(raw)));
  if (threadIdx.x == 0) {
    const uint32_t idesc = MODE == 0 ? idesc_f16(0, 0, N, 128) : idesc_f16(0, 1, N, 128);
    const uint64_t dA = umma_desc(sA, 0, 1024);
    const uint64_t dB = MODE == 0 ? umma_desc(sB, 0, 1024) : umma_desc(sB, 64 * 128, 1024);
    uint32_t ph = 0;
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int i = 0; i < per_commit; ++i) {
        if (MODE == 0)
          umma_f16(tmem, dA + (uint64_t)((i & 3) * 2), dB + (uint64_t)((i & 3) * 2), idesc, i > 0);
        else
          umma_f16_ts(tmem + 256, tmem + 384 + (uint32_t)((i & 3) * 8), dB + (uint64_t)((i & 3) * (2048 >> 4)), idesc, i > 0);
      }
      umma_commit(bar);
      mbar_wait(bar, ph);
      ph ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, int MODE>
void run(long long* d, int sms) {
  const int smem = 1024 + 128 * 128 + 256 * 128 + 64;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2][256];
  const int counts[2] = {8, 72};
  for (int c = 0; c < 2; ++c) {
    k<N, MODE><<<sms, 128, smem>>>(d, 20, counts[c]);
    cudaDeviceSynchronize();
    cudaMemcpy(h[c], d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  }
  // marginal cost per instruction = (t(72) - t(8)) / 64 on SM 0; the intercept is commit + barrier round trip
  const double per = (double)(h[1][0] - h[0][0]) / 64.0;
  printf("N=%3d A from %s: %6.1f clk per MMA (floor 128*N/256 = %3d), 8 MMAs + commit + wait = %lld clk, 72 -> %lld clk; err=%s\n", N, MODE == 0 ? "smem" : "TMEM",
         per, N / 2, h[0][0], h[1][0], cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 256);
  run<48, 0>(d, sms);
  run<64, 0>(d, sms);
  run<128, 0>(d, sms);
  run<256, 0>(d, sms);
  run<48, 1>(d, sms);
  run<64, 1>(d, sms);
  run<128, 1>(d, sms);
  run<256, 1>(d, sms);
  return 0;
}
