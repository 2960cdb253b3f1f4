// conv_in (Cin=8 -> C) and conv_out (C -> 8): k=3, pad=1 convolutions whose contraction (K=24) or output
// width (N=8) is too thin for a tensor-core tile.  They are HBM-bound (the wide activation is touched
// once), so they are written as direct convolutions that also do the layout change between the
// reference's [B, C, L] fp32 tensors and the internal channels-last bf16 activations.
#include "common.cuh"

namespace {

constexpr int POS = 32;  // positions per CTA

// y[b, l, co] = bias[co] + sum_{t,ci} w[co, ci, t] * x[b, ci, l + t - 1]
__global__ void conv_in_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ y,
                                   int Cin, int L, int Co) {
  extern __shared__ float sh[];
  float* sw = sh;                    // [3*Cin][Co]  (k-major so that co is contiguous)
  float* sx = sh + 3 * Cin * Co;     // [Cin][POS+2]
  const int b = blockIdx.y, l0 = blockIdx.x * POS;
  for (int i = threadIdx.x; i < Co * Cin * 3; i += blockDim.x) {
    const int co = i / (Cin * 3), rem = i % (Cin * 3), ci = rem / 3, t = rem % 3;
    sw[(t * Cin + ci) * Co + co] = w[i];
  }
  for (int i = threadIdx.x; i < Cin * (POS + 2); i += blockDim.x) {
    const int ci = i / (POS + 2), p = i % (POS + 2), l = l0 + p - 1;
    sx[i] = (l >= 0 && l < L) ? x[((long long)b * Cin + ci) * L + l] : 0.f;
  }
  __syncthreads();
  const int nv = Co >> 3;
  for (int i = threadIdx.x; i < POS * nv; i += blockDim.x) {
    const int p = i / nv, v = i % nv, l = l0 + p;
    if (l >= L) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias[v * 8 + j];
    for (int t = 0; t < 3; ++t)
      for (int ci = 0; ci < Cin; ++ci) {
        const float xv = sx[ci * (POS + 2) + p + t];
        const float* wr = sw + (t * Cin + ci) * Co + v * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wr[j], acc[j]);
      }
    store8(y + ((long long)b * L + l) * Co + v * 8, acc);
  }
}

// dw[co, ci, t] += sum_{b,l} dy[b,l,co] * x[b,ci,l+t-1] ; dbias[co] += sum dy
__global__ void conv_in_bwd_kernel(const bf16* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, float* __restrict__ dbias,
                                   int Cin, int L, int Co) {
  extern __shared__ float sh[];
  float* sd = sh;                  // [POS][Co]
  float* sx = sh + POS * Co;       // [Cin][POS+2]
  const int b = blockIdx.y, l0 = blockIdx.x * POS;
  for (int i = threadIdx.x; i < POS * Co; i += blockDim.x) {
    const int p = i / Co, co = i % Co, l = l0 + p;
    sd[i] = l < L ? __bfloat162float(dy[((long long)b * L + l) * Co + co]) : 0.f;
  }
  for (int i = threadIdx.x; i < Cin * (POS + 2); i += blockDim.x) {
    const int ci = i / (POS + 2), p = i % (POS + 2), l = l0 + p - 1;
    sx[i] = (l >= 0 && l < L) ? x[((long long)b * Cin + ci) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Co * Cin * 3; i += blockDim.x) {
    const int co = i / (Cin * 3), rem = i % (Cin * 3), ci = rem / 3, t = rem % 3;
    float acc = 0.f;
#pragma unroll 8
    for (int p = 0; p < POS; ++p) acc = fmaf(sd[p * Co + co], sx[ci * (POS + 2) + p + t], acc);
    atomicAdd(&dw[i], acc);
  }
  for (int co = threadIdx.x; co < Co; co += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < POS; ++p) acc += sd[p * Co + co];
    atomicAdd(&dbias[co], acc);
  }
}

// y[b, co, l] = bias[co] + sum_{t,c} w[co, c, t] * h[b, l+t-1, c]    one warp per position
__global__ void conv_out_fwd_kernel(const bf16* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                    int C, int L, int Cout) {
  extern __shared__ float sh[];  // sw[t][co][C]
  for (int i = threadIdx.x; i < Cout * C * 3; i += blockDim.x) {
    const int co = i / (C * 3), rem = i % (C * 3), c = rem / 3, t = rem % 3;
    sh[(t * Cout + co) * C + c] = w[i];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nv = C >> 3;
  for (int l = blockIdx.x * nwarps + warp; l < L; l += gridDim.x * nwarps) {
    float acc[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[co] = 0.f;
    for (int t = 0; t < 3; ++t) {
      const int ls = l + t - 1;
      if (ls < 0 || ls >= L) continue;
      const bf16* hr = h + ((long long)b * L + ls) * C;
      for (int v = lane; v < nv; v += 32) {
        float f[8];
        load8(hr + v * 8, f);
        for (int co = 0; co < Cout; ++co) {
          const float* wr = sh + (t * Cout + co) * C + v * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[co] = fmaf(f[j], wr[j], acc[co]);
        }
      }
    }
    for (int co = 0; co < Cout; ++co) {
      const float s = warp_sum(acc[co]);
      if (lane == 0) y[((long long)b * Cout + co) * L + l] = s + bias[co];
    }
  }
}

// dh[b, l, c] = sum_{t,co} dy[b, co, l - t + 1] * w[co, c, t]
__global__ void conv_out_bwd_dh_kernel(const float* __restrict__ dy, const float* __restrict__ w, bf16* __restrict__ dh, int C, int L, int Cout) {
  extern __shared__ float sh[];
  float* sw = sh;                      // [t][co][C]
  float* sd = sh + 3 * Cout * C;       // [Cout][POS+2]
  const int b = blockIdx.y, l0 = blockIdx.x * POS;
  for (int i = threadIdx.x; i < Cout * C * 3; i += blockDim.x) {
    const int co = i / (C * 3), rem = i % (C * 3), c = rem / 3, t = rem % 3;
    sw[(t * Cout + co) * C + c] = w[i];
  }
  for (int i = threadIdx.x; i < Cout * (POS + 2); i += blockDim.x) {
    const int co = i / (POS + 2), p = i % (POS + 2), l = l0 + p - 1;
    sd[i] = (l >= 0 && l < L) ? dy[((long long)b * Cout + co) * L + l] : 0.f;
  }
  __syncthreads();
  const int nv = C >> 3;
  for (int i = threadIdx.x; i < POS * nv; i += blockDim.x) {
    const int p = i / nv, v = i % nv, l = l0 + p;
    if (l >= L) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = 0; t < 3; ++t)
      for (int co = 0; co < Cout; ++co) {
        const float d = sd[co * (POS + 2) + (p + 1) - t + 1];  // position l - t + 1  -> local index (l - t + 1) - l0 + 1
        const float* wr = sw + (t * Cout + co) * C + v * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(d, wr[j], acc[j]);
      }
    store8(dh + ((long long)b * L + l) * C + v * 8, acc);
  }
}

// dw[co, c, t] += sum_{b,l} dy[b,co,l] * h[b,l+t-1,c] ; dbias[co] += sum dy
__global__ void conv_out_bwd_dw_kernel(const float* __restrict__ dy, const bf16* __restrict__ h, float* __restrict__ dw, float* __restrict__ dbias,
                                       int C, int L, int Cout) {
  extern __shared__ float sh[];
  float* shh = sh;                       // [POS+2][C]
  float* sd = sh + (POS + 2) * C;        // [Cout][POS]
  const int b = blockIdx.y, l0 = blockIdx.x * POS;
  for (int i = threadIdx.x; i < (POS + 2) * C; i += blockDim.x) {
    const int p = i / C, c = i % C, l = l0 + p - 1;
    shh[i] = (l >= 0 && l < L) ? __bfloat162float(h[((long long)b * L + l) * C + c]) : 0.f;
  }
  for (int i = threadIdx.x; i < Cout * POS; i += blockDim.x) {
    const int co = i / POS, p = i % POS, l = l0 + p;
    sd[i] = l < L ? dy[((long long)b * Cout + co) * L + l] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * C * 3; i += blockDim.x) {
    // iterate with c fastest for conflict-free shared reads; write to [co][c][t]
    const int co = i / (C * 3), rem = i % (C * 3), t = rem / C, c = rem % C;
    float acc = 0.f;
#pragma unroll 8
    for (int p = 0; p < POS; ++p) acc = fmaf(sd[co * POS + p], shh[(p + t) * C + c], acc);
    atomicAdd(&dw[((long long)co * C + c) * 3 + t], acc);
  }
  for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < POS; ++p) acc += sd[co * POS + p];
    atomicAdd(&dbias[co], acc);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) PT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return PT_OK;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int pt_conv_in_fwd(const float* x_ncl, const float* w, const float* bias, void* y, int B, int Cin, int L, int Co, void* stream) {
  PT_REQUIRE(B > 0 && Cin > 0 && Cin <= 16 && L > 0 && Co % 8 == 0, "conv_in_fwd: Cin=%d Co=%d", Cin, Co);
  const size_t smem = sizeof(float) * (3 * Cin * Co + Cin * (POS + 2));
  PT_REQUIRE(smem <= 200 * 1024, "conv_in_fwd: Co=%d too large", Co);
  if (int r = set_smem(conv_in_fwd_kernel, smem)) return r;
  conv_in_fwd_kernel<<<dim3((L + POS - 1) / POS, B), 256, smem, ST>>>(x_ncl, w, bias, (bf16*)y, Cin, L, Co);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
extern "C" int pt_conv_in_bwd(const void* dy, const float* x_ncl, float* dw, float* dbias, int B, int Cin, int L, int Co, void* stream) {
  PT_REQUIRE(B > 0 && Cin > 0 && Cin <= 16 && L > 0 && Co % 8 == 0, "conv_in_bwd: Cin=%d Co=%d", Cin, Co);
  const size_t smem = sizeof(float) * (POS * Co + Cin * (POS + 2));
  PT_REQUIRE(smem <= 200 * 1024, "conv_in_bwd: Co=%d too large", Co);
  if (int r = set_smem(conv_in_bwd_kernel, smem)) return r;
  conv_in_bwd_kernel<<<dim3((L + POS - 1) / POS, B), 256, smem, ST>>>((const bf16*)dy, x_ncl, dw, dbias, Cin, L, Co);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
extern "C" int pt_conv_out_fwd(const void* h, const float* w, const float* bias, float* y_ncl, int B, int C, int L, int Cout, void* stream) {
  PT_REQUIRE(B > 0 && C % 8 == 0 && L > 0 && Cout > 0 && Cout <= 8, "conv_out_fwd: C=%d Cout=%d", C, Cout);
  const size_t smem = sizeof(float) * 3 * Cout * C;
  PT_REQUIRE(smem <= 200 * 1024, "conv_out_fwd: C=%d too large", C);
  if (int r = set_smem(conv_out_fwd_kernel, smem)) return r;
  int gx = (L + 7) / 8;
  const int cap = (4 * pt_num_sms() + B - 1) / B;
  if (gx > cap) gx = cap;
  conv_out_fwd_kernel<<<dim3(gx, B), 256, smem, ST>>>((const bf16*)h, w, bias, y_ncl, C, L, Cout);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
extern "C" int pt_conv_out_bwd(const float* dy_ncl, const void* h, const float* w, void* dh, float* dw, float* dbias, int B, int C, int L,
                               int Cout, void* stream) {
  PT_REQUIRE(B > 0 && C % 8 == 0 && L > 0 && Cout > 0 && Cout <= 8, "conv_out_bwd: C=%d Cout=%d", C, Cout);
  const size_t smem1 = sizeof(float) * (3 * Cout * C + Cout * (POS + 2));
  const size_t smem2 = sizeof(float) * ((POS + 2) * C + Cout * POS);
  PT_REQUIRE(smem1 <= 200 * 1024 && smem2 <= 200 * 1024, "conv_out_bwd: C=%d too large", C);
  if (int r = set_smem(conv_out_bwd_dh_kernel, smem1)) return r;
  if (int r = set_smem(conv_out_bwd_dw_kernel, smem2)) return r;
  conv_out_bwd_dh_kernel<<<dim3((L + POS - 1) / POS, B), 256, smem1, ST>>>(dy_ncl, w, (bf16*)dh, C, L, Cout);
  PT_LAUNCH_CHECK();
  conv_out_bwd_dw_kernel<<<dim3((L + POS - 1) / POS, B), 256, smem2, ST>>>(dy_ncl, (const bf16*)h, dw, dbias, C, L, Cout);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
