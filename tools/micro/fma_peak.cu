// FP32 FMA-pipe peak of the device (the roofline denominator of the exact-fp32 RVQ quantiser; MEASURED_PEAKS.json has no such entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o tools/micro/fma_peak tools/micro/fma_peak.cu
// Every thread runs 16 independent fmaf chains; 2048 threads per SM; prints one JSON line (TFLOP/s, best of 5).
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void __launch_bounds__(1024, 2) fma_kernel(float* out, int iters, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  if (s == 12345.678f) out[0] = s;   // never true: keeps the chains alive
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 1 << 16;
  const int blocks = sms * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fma_kernel<<<blocks, 1024>>>(out, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    fma_kernel<<<blocks, 1024>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * (double)iters * 1024.0 * blocks;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  printf("{\"fp32_fma_tflops\": %.2f, \"sms\": %d, \"how\": \"16 independent fmaf chains per thread, 2048 threads per SM, 65536 iterations, best of 5 (CUDA events)\"}\n", best, sms);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
