// GroupNorm (+SiLU) and LayerNorm, forward and backward, channels-last bf16 activations, fp32 math.
// HBM-bound: 16-byte vector accesses, one pass for statistics, one for the apply.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics.  grid = (column blocks of 256 channels, row chunks, batch).  A warp reads 32 consecutive
// 16-byte vectors of a row; the 8 warps take rows w, w+8, ... four at a time, so per-channel partials stay in
// registers; they are folded into per-group sums through shared memory and one atomicAdd per (CTA, group).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sums, int L, int C, int G,
                                                       int rows_per_cta) {
  extern __shared__ float sh[];  // [2*G]
  const int b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(L, r0 + rows_per_cta);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  if (c < C) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const bf16* xb = x + ((long long)b * L) * C + c;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(xb + (long long)(r + 8 * u) * C, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += f[u][j];
          q[j] += f[u][j] * f[u][j];
        }
    }
    for (; r < r1; r += 8) {
      float f[8];
      load8(xb + (long long)r * C, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] += f[j] * f[j];
      }
    }
    const int cpg = C / G;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      atomicAdd(&sh[2 * g], s[j]);
      atomicAdd(&sh[2 * g + 1], q[j]);
    }
  }
  __syncthreads();
  const int cpg = C / G;
  const int g0 = (blockIdx.x * 256) / cpg, g1 = min(G - 1, (min(C, blockIdx.x * 256 + 256) - 1) / cpg);
  for (int i = 2 * g0 + threadIdx.x; i <= 2 * g1 + 1; i += blockDim.x) atomicAdd(&sums[(long long)b * 2 * G + i], sh[i]);
}

__global__ void gn_finalize_kernel(float* __restrict__ stats, int n, float inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float mean = stats[2 * i] * inv_count;
    const float var = fmaxf(stats[2 * i + 1] * inv_count - mean * mean, 0.f);
    stats[2 * i] = mean;
    stats[2 * i + 1] = rsqrtf(var + eps);
  }
}

// y = act(x * a[c] + b[c]) with a = rstd*gamma, b = beta - mean*a held in shared memory per batch element.
__global__ void gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, bf16* __restrict__ y, int L, int C, int G, int rows_per_cta, int act) {
  extern __shared__ float sh[];  // a[C], b[C]
  float* sa = sh;
  float* sb = sh + C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1];
    const float a = rstd * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(L, r0 + rows_per_cta);
  const bf16* xb = x + ((long long)b * L + r0) * C;
  bf16* yb = y + ((long long)b * L + r0) * C;
  const int total = (r1 - r0) * nvec;   // the chunk is contiguous: vector i covers elements [8i, 8i+8)
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) load8(xb + (long long)i * 8, f[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) {
        const int v = i % nvec;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float z = f[u][j] * sa[v * 8 + j] + sb[v * 8 + j];
          f[u][j] = act ? silu_f(z) : z;
        }
        store8(yb + (long long)i * 8, f[u]);
      }
    }
  }
}

// backward pass 1: per-channel sum(dz), sum(dz*xhat) -> dgamma/dbeta (global atomics) and
// per-(b,g) s1 = sum(dz*gamma), s2 = sum(dz*gamma*xhat) -> scratch.  Same tiling as gn_stats_kernel.
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ scratch,
                                                            int L, int C, int G, int rows_per_cta, int act) {
  extern __shared__ float sh[];  // [2*G] group sums, then [8][256][2] per-warp channel partials
  float* chp = sh + 2 * G;
  const int b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(L, r0 + rows_per_cta);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh[i] = 0.f;
  const int cpg = C / G;
  float sdz[8], sdzx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sdz[j] = sdzx[j] = 0.f;
  if (c < C) {
    float mean[8], rstd[8], gm[8], bt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      mean[j] = stats[((long long)b * G + g) * 2];
      rstd[j] = stats[((long long)b * G + g) * 2 + 1];
      gm[j] = gamma[c + j];
      bt[j] = beta[c + j];
    }
    const long long base = ((long long)b * L) * C + c;
    int r = r0 + warp;
    for (; r + 8 < r1; r += 16) {
      float fx[2][8], fd[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        load8(x + base + (long long)(r + 8 * u) * C, fx[u]);
        load8(dy + base + (long long)(r + 8 * u) * C, fd[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (fx[u][j] - mean[j]) * rstd[j];
          float dz = fd[u][j];
          if (act) dz *= silu_grad_f(xh * gm[j] + bt[j]);
          sdz[j] += dz;
          sdzx[j] += dz * xh;
        }
    }
    for (; r < r1; r += 8) {
      float fx[8], fd[8];
      load8(x + base + (long long)r * C, fx);
      load8(dy + base + (long long)r * C, fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fx[j] - mean[j]) * rstd[j];
        float dz = fd[j];
        if (act) dz *= silu_grad_f(xh * gm[j] + bt[j]);
        sdz[j] += dz;
        sdzx[j] += dz * xh;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    chp[(warp * 256 + lane * 8 + j) * 2] = sdz[j];
    chp[(warp * 256 + lane * 8 + j) * 2 + 1] = sdzx[j];
  }
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < C) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += chp[(w * 256 + threadIdx.x) * 2];
      bsum += chp[(w * 256 + threadIdx.x) * 2 + 1];
    }
    atomicAdd(&dbeta[cc], a);
    atomicAdd(&dgamma[cc], bsum);
    const int g = cc / cpg;
    const float gmc = gamma[cc];
    atomicAdd(&sh[2 * g], a * gmc);
    atomicAdd(&sh[2 * g + 1], bsum * gmc);
  }
  __syncthreads();
  const int g0 = (blockIdx.x * 256) / cpg, g1 = min(G - 1, (min(C, blockIdx.x * 256 + 256) - 1) / cpg);
  for (int i = 2 * g0 + threadIdx.x; i <= 2 * g1 + 1; i += blockDim.x) atomicAdd(&scratch[(long long)b * 2 * G + i], sh[i]);
}

// backward pass 2: dx = rstd * (dz*gamma - s1/n - xhat * s2/n)
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scratch,
                                    bf16* __restrict__ dx, int L, int C, int G, int rows_per_cta, int act) {
  extern __shared__ float sh[];  // per channel: mean, rstd, gamma, beta, k1 (= s1/n), k2 (= s2/n)
  float* s_mean = sh;
  float* s_rstd = sh + C;
  float* s_g = sh + 2 * C;
  float* s_b = sh + 3 * C;
  float* s_k1 = sh + 4 * C;
  float* s_k2 = sh + 5 * C;
  const int b = blockIdx.y;
  const int cpg = C / G;
  const float inv_n = 1.f / ((float)cpg * (float)L);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    s_mean[c] = stats[((long long)b * G + g) * 2];
    s_rstd[c] = stats[((long long)b * G + g) * 2 + 1];
    s_g[c] = gamma[c];
    s_b[c] = beta[c];
    s_k1[c] = scratch[((long long)b * G + g) * 2] * inv_n;
    s_k2[c] = scratch[((long long)b * G + g) * 2 + 1] * inv_n;
  }
  __syncthreads();
  const int nvec = C >> 3;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(L, r0 + rows_per_cta);
  const long long base = ((long long)b * L + r0) * C;
  const int total = (r1 - r0) * nvec;
  for (int i0 = threadIdx.x; i0 < total; i0 += 2 * blockDim.x) {
    float fx[2][8], fd[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) {
        load8(x + base + (long long)i * 8, fx[u]);
        load8(dy + base + (long long)i * 8, fd[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < total) {
        const int v = i % nvec;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = v * 8 + j;
          const float xh = (fx[u][j] - s_mean[c]) * s_rstd[c];
          float dz = fd[u][j];
          if (act) dz *= silu_grad_f(xh * s_g[c] + s_b[c]);
          fd[u][j] = s_rstd[c] * (dz * s_g[c] - s_k1[c] - xh * s_k2[c]);
        }
        store8(dx + base + (long long)i * 8, fd[u]);
      }
    }
  }
}

struct GnGeom {
  int rpp, threads, rows_per_cta, chunks;   // elementwise passes: grid (chunks, B)
  int cblocks, red_rows, red_chunks;         // reduction passes: grid (cblocks, red_chunks, B)
};
int gn_geom(int B, int L, int C, GnGeom* g) {
  const int nvec = C / 8;
  PT_REQUIRE(C % 8 == 0 && nvec <= 1024, "groupnorm: C=%d must be a multiple of 8 and <= 8192", C);
  g->rpp = nvec >= 256 ? 1 : 256 / nvec;
  g->threads = (nvec * g->rpp + 31) / 32 * 32;
  int want = (4 * pt_num_sms() + B - 1) / B;  // chunks per batch element
  int rows = (L + want - 1) / want;
  rows = (rows + g->rpp - 1) / g->rpp * g->rpp;
  if (rows < 4 * g->rpp) rows = 4 * g->rpp;
  g->rows_per_cta = rows;
  g->chunks = (L + rows - 1) / rows;
  g->cblocks = (C + 255) / 256;
  int rwant = (8 * pt_num_sms() + B * g->cblocks - 1) / (B * g->cblocks);
  int rr = (L + rwant - 1) / rwant;
  if (rr < 32) rr = 32;
  g->red_rows = rr;
  g->red_chunks = (L + rr - 1) / rr;
  return PT_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C <= 2048).
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;

template <int NV>
__global__ void ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              bf16* __restrict__ y, float* __restrict__ rowstats, long long M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = C >> 3;
  float f[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      load8(x + row * C + v * 8, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * __ldg(gamma + v * 8 + j) + __ldg(beta + v * 8 + j);
      store8(y + row * C + v * 8, o);
    }
  }
  if (lane == 0) {
    rowstats[2 * row] = mean;
    rowstats[2 * row + 1] = rstd;
  }
}

// Each warp walks rows w, w + nwarps_total, ...; per-lane dgamma/dbeta partials stay in registers and are
// flushed once through shared memory + global atomics.
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ rowstats,
                              const float* __restrict__ gamma, const bf16* __restrict__ dx_add, bf16* __restrict__ dx,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C) {
  extern __shared__ float sh[];  // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += wstride) {
    const float mean = rowstats[2 * row], rstd = rowstats[2 * row + 1];
    float xh[NV][8], dg[NV][8];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float fx[8], fd[8];
        load8(x + row * C + v * 8, fx);
        load8(dy + row * C + v * 8, fd);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[i][j] = (fx[j] - mean) * rstd;
          ag[i][j] += fd[j] * xh[i][j];
          ab[i][j] += fd[j];
          dg[i][j] = fd[j] * __ldg(gamma + v * 8 + j);
          c1 += dg[i][j];
          c2 += dg[i][j] * xh[i][j];
        }
      }
    }
    c1 = warp_sum(c1) / (float)C;
    c2 = warp_sum(c2) / (float)C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (dg[i][j] - c1 - xh[i][j] * c2);
        if (dx_add) {
          float t[8];
          load8(dx_add + row * C + v * 8, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j];
        }
        store8(dx + row * C + v * 8, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sh[v * 8 + j], ag[i][j]);
        atomicAdd(&sh[C + v * 8 + j], ab[i][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dgamma[i], sh[i]);
    atomicAdd(&dbeta[i], sh[C + i]);
  }
}

}  // namespace

extern "C" int pt_groupnorm_stats(const void* x, float* stats, int B, int L, int C, int G, float eps, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_stats: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  PT_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * B * G, st));
  gn_stats_kernel<<<dim3(g.cblocks, g.red_chunks, B), 256, 2 * G * sizeof(float), st>>>((const bf16*)x, stats, L, C, G, g.red_rows);
  PT_LAUNCH_CHECK();
  const int n = B * G;
  gn_finalize_kernel<<<(n + 127) / 128, 128, 0, st>>>(stats, n, 1.f / ((float)(C / G) * (float)L), eps);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_groupnorm_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y, int B, int L, int C,
                                  int G, int act, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_apply: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  gn_apply_kernel<<<dim3(g.chunks, B), 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>((const bf16*)x, stats, gamma, beta, (bf16*)y, L, C, G,
                                                                                          g.rows_per_cta, act);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_groupnorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, void* dx,
                                float* dgamma, float* dbeta, float* scratch, int B, int L, int C, int G, int act, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_bwd: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  PT_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * B * G, st));
  gn_bwd_reduce_kernel<<<dim3(g.cblocks, g.red_chunks, B), 256, (2 * G + 8 * 256 * 2) * sizeof(float), st>>>(
      (const bf16*)dy, (const bf16*)x, stats, gamma, beta, dgamma, dbeta, scratch, L, C, G, g.red_rows, act);
  PT_LAUNCH_CHECK();
  PT_REQUIRE(6 * C * sizeof(float) <= 160 * 1024, "groupnorm_bwd: C=%d too large", C);
  static bool attr_set = false;
  if (!attr_set) {
    PT_CUDA_OK(cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  gn_bwd_apply_kernel<<<dim3(g.chunks, B), 256, 6 * C * sizeof(float), st>>>((const bf16*)dy, (const bf16*)x, stats, gamma, beta, scratch,
                                                                            (bf16*)dx, L, C, G, g.rows_per_cta, act);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* rowstats, int64_t M, int C, float eps,
                                void* stream) {
  PT_REQUIRE(M > 0 && C % 8 == 0 && C / 8 <= 32 * LN_MAXV, "layernorm_fwd: M=%lld C=%d", (long long)M, C);
  const int wpb = 8;
  const int nv = (C / 8 + 31) / 32;
#define LN_FWD(NV_)                                                                                                                  \
  case NV_:                                                                                                                          \
    ln_fwd_kernel<NV_><<<(unsigned)((M + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>((const bf16*)x, gamma, beta, (bf16*)y, \
                                                                                               rowstats, M, C, eps);                 \
    break;
  switch (nv) {
    LN_FWD(1) LN_FWD(2) LN_FWD(3) LN_FWD(4) LN_FWD(5) LN_FWD(6) LN_FWD(7) LN_FWD(8)
  }
#undef LN_FWD
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_layernorm_bwd(const void* dy, const void* x, const float* rowstats, const float* gamma, const void* dx_add, void* dx,
                                float* dgamma, float* dbeta, int64_t M, int C, void* stream) {
  PT_REQUIRE(M > 0 && C % 8 == 0 && C / 8 <= 32 * LN_MAXV, "layernorm_bwd: M=%lld C=%d", (long long)M, C);
  const int wpb = 8;
  long long blocks = (M + wpb - 1) / wpb;
  const long long cap = 2ll * pt_num_sms();
  if (blocks > cap) blocks = cap;
  const int nv = (C / 8 + 31) / 32;
#define LN_BWD(NV_)                                                                                                            \
  case NV_:                                                                                                                    \
    ln_bwd_kernel<NV_><<<(unsigned)blocks, wpb * 32, 2 * C * sizeof(float), (cudaStream_t)stream>>>(                            \
        (const bf16*)dy, (const bf16*)x, rowstats, gamma, (const bf16*)dx_add, (bf16*)dx, dgamma, dbeta, M, C);                \
    break;
  switch (nv) {
    LN_BWD(1) LN_BWD(2) LN_BWD(3) LN_BWD(4) LN_BWD(5) LN_BWD(6) LN_BWD(7) LN_BWD(8)
  }
#undef LN_BWD
  PT_LAUNCH_CHECK();
  return PT_OK;
}
