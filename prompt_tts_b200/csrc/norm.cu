// GroupNorm (+SiLU) and LayerNorm, forward and backward, channels-last bf16 activations, fp32 math.
// HBM-bound: 16-byte vector accesses, one pass for statistics, one for the apply.
#include <stdlib.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// GroupNorm.  All four kernels share one mapping: grid = (row chunks, batch); a CTA has nvec * rpp threads
// (nvec = C / 8 sixteen-byte vectors per row); thread t owns vector t % nvec of rows r0 + t / nvec, + rpp, ...
// Consecutive threads touch consecutive 16-byte vectors (also across a row boundary), so every warp access is one
// contiguous 512-byte segment, and everything that depends only on the channel -- folded normalisation coefficients,
// per-channel partial sums -- lives in registers for the whole row loop.  Reductions go registers -> shared memory
// (one slot per thread, no atomics) -> per-channel -> per-group, then one global atomic per value per CTA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_fast(float z) {   // one MUFU op (exp + rcp would be two: these kernels are MUFU-bound otherwise)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

struct GnMap {
  int nvec, rpp, threads, v, rsub;
};
__device__ __forceinline__ GnMap gn_map(int C) {
  GnMap m;
  m.nvec = C >> 3;
  m.threads = blockDim.x;
  m.rpp = m.threads / m.nvec;
  m.v = threadIdx.x % m.nvec;
  m.rsub = threadIdx.x / m.nvec;
  return m;
}

// per-channel (a, b) with  gn(x) * gamma + beta = x * a + b  for this thread's 8 channels.
// fast (host-checked: gamma/beta 16-byte aligned, C/G >= 8 so the 8 channels span at most two groups): six vector loads
// instead of 32 scalar ones -- this prologue is on the critical path of every CTA.
template <bool FAST>
__device__ __forceinline__ void gn_coeffs(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta, int b,
                                          int c0, int cpg, int G, float* a1, float* b1, float* rs, float* ms) {
  float gm[8], bt[8], mean[8], rstd[8];
  if constexpr (FAST) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), e1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const int ga = c0 / cpg, gb = (c0 + 7) / cpg;
    const float2 sa = __ldg(reinterpret_cast<const float2*>(stats) + (long long)b * G + ga);
    const float2 sb = __ldg(reinterpret_cast<const float2*>(stats) + (long long)b * G + gb);
    const int nb = (ga + 1) * cpg - c0;   // channels of this vector that belong to the first group
    gm[0] = g0.x, gm[1] = g0.y, gm[2] = g0.z, gm[3] = g0.w, gm[4] = g1.x, gm[5] = g1.y, gm[6] = g1.z, gm[7] = g1.w;
    bt[0] = e0.x, bt[1] = e0.y, bt[2] = e0.z, bt[3] = e0.w, bt[4] = e1.x, bt[5] = e1.y, bt[6] = e1.z, bt[7] = e1.w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = j < nb ? sa.x : sb.x;
      rstd[j] = j < nb ? sa.y : sb.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cpg;
      mean[j] = stats[((long long)b * G + g) * 2];
      rstd[j] = stats[((long long)b * G + g) * 2 + 1];
      gm[j] = gamma[c0 + j];
      bt[j] = beta[c0 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = rstd[j] * gm[j];
    b1[j] = bt[j] - mean[j] * rstd[j] * gm[j];
    if (rs) {
      rs[j] = rstd[j];
      ms[j] = -mean[j] * rstd[j];
    }
  }
}

// All four kernels stream rows in batches of U per thread: the 16-byte loads of a batch are issued back to back as raw bf16
// (4 registers each, converted when consumed) and the per-channel coefficients are fetched after the first batch is already in
// flight; several CTAs per SM cover each other's round trips.  A whole tensor here is only 100-600 KB per SM, i.e. a few memory round
// trips: what matters is that every round trip carries as many bytes as possible and that nothing serialises in front of it.
__global__ void __launch_bounds__(320) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sums, int L, int C, int G, int rows_per_cta) {
  extern __shared__ float sh[];  // [threads][16] partials, reused as [C][2]
  const GnMap m = gn_map(C);
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(L, r0 + rows_per_cta);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const bf16* xb = x + ((long long)b * L) * C + m.v * 8;
  int r = r0 + m.rsub;
  for (; r + 3 * m.rpp < r1; r += 4 * m.rpp) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load8(xb + (long long)(r + u * m.rpp) * C, f[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[u][j];
        q[j] = fmaf(f[u][j], f[u][j], q[j]);
      }
  }
  for (; r < r1; r += m.rpp) {
    float f[8];
    load8(xb + (long long)r * C, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      q[j] = fmaf(f[j], f[j], q[j]);
    }
  }
  // fold the rpp row-phases: slot [rsub][channel][2]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[((m.rsub * C) + m.v * 8 + j) * 2] = s[j];
    sh[((m.rsub * C) + m.v * 8 + j) * 2 + 1] = q[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += m.threads) {
    float a = sh[c * 2], d = sh[c * 2 + 1];
    for (int k = 1; k < m.rpp; ++k) {
      a += sh[(k * C + c) * 2];
      d += sh[(k * C + c) * 2 + 1];
    }
    sh[c * 2] = a;
    sh[c * 2 + 1] = d;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += m.threads) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cpg; ++k) {
      a += sh[(g * cpg + k) * 2];
      d += sh[(g * cpg + k) * 2 + 1];
    }
    atomicAdd(&sums[((long long)b * G + g) * 2], a);
    atomicAdd(&sums[((long long)b * G + g) * 2 + 1], d);
  }
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

// y = act(x * a[c] + b[c]) with a = rstd*gamma, b = beta - mean*a.  Rows are streamed in batches of U per thread: the 16-byte
// loads of a batch are issued back to back as raw bf16 (4 registers each, converted when consumed) and the per-channel
// coefficients are fetched after the first batch is already in flight -- a tensor here is only 100-600 KB per SM, a few memory
// round trips, so what matters is that each round trip carries many bytes and nothing serialises in front of it (9.2 vs 12.3 us
// at 32 x 752 x 320).  The same restructuring made the three reducing kernels slower (their flush dominates) and was not kept.
template <int U, bool FAST>
__global__ void __launch_bounds__(320) gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, bf16* __restrict__ y, int L, int C, int G, int rows_per_cta,
                                                       int act) {
  const GnMap m = gn_map(C);
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(L, r0 + rows_per_cta);
  const long long base = ((long long)b * L) * C + m.v * 8;
  const int step = U * m.rpp;
  int r = r0 + m.rsub;
  bf16x8 cur[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (r + u * m.rpp < r1) cur[u] = load8raw(x + base + (long long)(r + u * m.rpp) * C);
  float a1[8], b1[8];
  gn_coeffs<FAST>(stats, gamma, beta, b, m.v * 8, C / G, G, a1, b1, nullptr, nullptr);
  while (r < r1) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * m.rpp < r1) {
        float f[8];
        unpack8(cur[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(f[j], a1[j], b1[j]);
          f[j] = act ? z * sigmoid_fast(z) : z;
        }
        store8(y + base + (long long)(r + u * m.rpp) * C, f);
      }
    }
    r += step;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (r + u * m.rpp < r1) cur[u] = load8raw(x + base + (long long)(r + u * m.rpp) * C);
  }
}

// backward pass 1: per-channel sum(dz), sum(dz*xhat) -> dgamma/dbeta (global atomics) and
// per-(b,g) s1 = sum(dz*gamma), s2 = sum(dz*gamma*xhat) -> scratch.
template <bool FAST>
__global__ void __launch_bounds__(320) gn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ scratch,
                                                            int L, int C, int G, int rows_per_cta, int act) {
  extern __shared__ float sh[];  // [rpp][C][2]
  const GnMap m = gn_map(C);
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(L, r0 + rows_per_cta);
  const int cpg = C / G;
  float a1[8], b1[8], rs[8], ms[8], sdz[8], sdzx[8];
  gn_coeffs<FAST>(stats, gamma, beta, b, m.v * 8, cpg, G, a1, b1, rs, ms);
#pragma unroll
  for (int j = 0; j < 8; ++j) sdz[j] = sdzx[j] = 0.f;
  const long long base = ((long long)b * L) * C + m.v * 8;
  auto body = [&](const float* fx, const float* fd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dz = fd[j];
      if (act) {
        const float z = fmaf(fx[j], a1[j], b1[j]);
        const float sg = sigmoid_fast(z);
        dz *= sg * fmaf(z, 1.f - sg, 1.f);
      }
      sdz[j] += dz;
      sdzx[j] = fmaf(dz, fmaf(fx[j], rs[j], ms[j]), sdzx[j]);
    }
  };
  int r = r0 + m.rsub;
  for (; r + m.rpp < r1; r += 2 * m.rpp) {
    float fx[2][8], fd[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      load8(x + base + (long long)(r + u * m.rpp) * C, fx[u]);
      load8(dy + base + (long long)(r + u * m.rpp) * C, fd[u]);
    }
    body(fx[0], fd[0]);
    body(fx[1], fd[1]);
  }
  for (; r < r1; r += m.rpp) {
    float fx[8], fd[8];
    load8(x + base + (long long)r * C, fx);
    load8(dy + base + (long long)r * C, fd);
    body(fx, fd);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[((m.rsub * C) + m.v * 8 + j) * 2] = sdz[j];
    sh[((m.rsub * C) + m.v * 8 + j) * 2 + 1] = sdzx[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += m.threads) {
    float a = sh[c * 2], d = sh[c * 2 + 1];
    for (int k = 1; k < m.rpp; ++k) {
      a += sh[(k * C + c) * 2];
      d += sh[(k * C + c) * 2 + 1];
    }
    atomicAdd(&dbeta[c], a);
    atomicAdd(&dgamma[c], d);
    const float gmc = gamma[c];
    sh[c * 2] = a * gmc;
    sh[c * 2 + 1] = d * gmc;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += m.threads) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cpg; ++k) {
      a += sh[(g * cpg + k) * 2];
      d += sh[(g * cpg + k) * 2 + 1];
    }
    atomicAdd(&scratch[((long long)b * G + g) * 2], a);
    atomicAdd(&scratch[((long long)b * G + g) * 2 + 1], d);
  }
}

// backward pass 2: dx = rstd * (dz*gamma - s1/n - xhat * s2/n)  =  dz * a1 + x * c2 + c3   (per-channel coefficients in registers)
template <bool FAST>
__global__ void __launch_bounds__(320) gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ scratch, const bf16* __restrict__ dx_add, bf16* dx, int L,
                                                           int C, int G, int rows_per_cta, int act) {
  const GnMap m = gn_map(C);
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(L, r0 + rows_per_cta);
  const int cpg = C / G;
  const float inv_n = 1.f / ((float)cpg * (float)L);
  float a1[8], b1[8], c2[8], c3[8];
  {
    float rs[8], ms[8];
    gn_coeffs<FAST>(stats, gamma, beta, b, m.v * 8, cpg, G, a1, b1, rs, ms);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (m.v * 8 + j) / cpg;
      const float k1 = scratch[((long long)b * G + g) * 2] * inv_n, k2 = scratch[((long long)b * G + g) * 2 + 1] * inv_n;
      c2[j] = -rs[j] * rs[j] * k2;            // -rstd^2 k2
      c3[j] = -rs[j] * k1 - ms[j] * rs[j] * k2;   // -rstd k1 + mean rstd^2 k2   (ms = -mean rstd)
    }
  }
  const long long base = ((long long)b * L) * C + m.v * 8;
  auto body = [&](const float* fx, float* fd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float dz = fd[j];
      if (act) {
        const float z = fmaf(fx[j], a1[j], b1[j]);
        const float sg = sigmoid_fast(z);
        dz *= sg * fmaf(z, 1.f - sg, 1.f);
      }
      fd[j] = fmaf(dz, a1[j], fmaf(fx[j], c2[j], c3[j]));
    }
  };
  int r = r0 + m.rsub;
  for (; r + m.rpp < r1; r += 2 * m.rpp) {
    float fx[2][8], fd[2][8];
    bf16x8 ta[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      load8(x + base + (long long)(r + u * m.rpp) * C, fx[u]);
      load8(dy + base + (long long)(r + u * m.rpp) * C, fd[u]);
      // fused accumulation of the gradient that reached x through the other branch.  dx may alias dx_add, so these loads must be
      // issued here, with the others and before any store of this iteration (each element is read and written by the same
      // thread): left next to their use they cannot be hoisted above the previous row's store and cost a second round trip
      if (dx_add) ta[u] = load8raw(dx_add + base + (long long)(r + u * m.rpp) * C);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      body(fx[u], fd[u]);
      if (dx_add) {
        float t[8];
        unpack8(ta[u], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) fd[u][j] += t[j];
      }
      store8(dx + base + (long long)(r + u * m.rpp) * C, fd[u]);
    }
  }
  for (; r < r1; r += m.rpp) {
    float fx[8], fd[8];
    bf16x8 ta;
    load8(x + base + (long long)r * C, fx);
    load8(dy + base + (long long)r * C, fd);
    if (dx_add) ta = load8raw(dx_add + base + (long long)r * C);
    body(fx, fd);
    if (dx_add) {
      float t[8];
      unpack8(ta, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) fd[j] += t[j];
    }
    store8(dx + base + (long long)r * C, fd);
  }
}

constexpr int GN_U_APPLY = 8;

struct GnGeom {
  int threads, rpp;
  int rows_apply, chunks_apply;   // forward apply: one batch of GN_U_APPLY rows per thread
  int rows_bapply, chunks_bapply; // backward apply
  int rows_red, chunks_red;       // reduction passes (fewer CTAs: each flushes 2C + 2G atomics)
};
int gn_geom(int B, int L, int C, GnGeom* g) {
  const int nvec = C / 8;
  PT_REQUIRE(C % 8 == 0 && nvec >= 1 && nvec <= 320, "groupnorm: C=%d must be a multiple of 8 and <= 2560", C);
  g->rpp = nvec >= 160 ? 1 : (256 / nvec > 0 ? 256 / nvec : 1);
  if (g->rpp > L) g->rpp = L;
  g->threads = nvec * g->rpp;
  auto pick = [&](int target_ctas, int min_rows_per_thread, int* rows, int* chunks) {
    int want = (target_ctas + B - 1) / B;
    int r = (L + want - 1) / want;
    if (r < min_rows_per_thread * g->rpp) r = min_rows_per_thread * g->rpp;
    r = (r + g->rpp - 1) / g->rpp * g->rpp;
    *rows = r;
    *chunks = (L + r - 1) / r;
  };
  pick(16 * pt_num_sms(), GN_U_APPLY, &g->rows_apply, &g->chunks_apply);
  pick(6 * pt_num_sms(), 4, &g->rows_bapply, &g->chunks_bapply);
  pick(3 * pt_num_sms(), 8, &g->rows_red, &g->chunks_red);
  return PT_OK;
}
// the vectorised coefficient fetch needs 16-byte aligned gamma / beta, 8-byte aligned stats and groups of at least 8 channels
int gn_fast(const float* stats, const float* gamma, const float* beta, int C, int G) {
  return ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0 &&
         C / G >= 8;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm forward with the sample resident in shared memory: a cluster of GN_CL CTAs owns one batch element; CTA k streams its
// slice of rows ([L / GN_CL rows] x C) from HBM ONCE, stashing the raw vectors in shared memory while it reduces them, exchanges
// the per-group partial sums with its peers through distributed shared memory, and normalises from the resident copy: 1 read + 1
// write of the tensor instead of 2 reads + 1 write (statistics kernel + apply kernel).  Same thread mapping and arithmetic as the
// streaming kernels above.  Measured cold (tools/gn_probe.py, B = 32): 14.2 vs 16.5 us at 752 x 320, 14.7 vs 18.1 at 376 x 640,
// 17.0 vs 19.6 at 188 x 1280 -- as long as a CTA's slice is small enough for two CTAs per SM.  Beyond that (one CTA per SM) only
// ~8 clusters of 8 are resident at a time and the kernel runs in 3-4 waves (752 x 640: 41.8 vs 24.5 us), which is also why the
// backward pass (x AND dy resident: twice the bytes) stays on the streaming pair: a resident-sample backward kernel measured 59 vs
// 33 us at 752 x 320.  (Loading the slice with cp.async.bulk instead of vector loads made no difference either way.)
// ------------------------------------------------------------------------------------------------
constexpr int GN_CL = 8;

__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// a float at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ float gn_ld_peer(const float* p, uint32_t rank) {
  uint32_t a;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(gn_smem_u32(p)), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
// per-channel coefficients of this thread's 8 channels, statistics (mean, rstd per group) in shared memory
__device__ __forceinline__ void gn_coeffs_sm(const float* sstat, const float* __restrict__ gamma, const float* __restrict__ beta, int c0, int cpg,
                                             bool vec, float* a1, float* b1, float* rs, float* ms) {
  float gm[8], bt[8];
  if (vec) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), e1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    gm[0] = g0.x, gm[1] = g0.y, gm[2] = g0.z, gm[3] = g0.w, gm[4] = g1.x, gm[5] = g1.y, gm[6] = g1.z, gm[7] = g1.w;
    bt[0] = e0.x, bt[1] = e0.y, bt[2] = e0.z, bt[3] = e0.w, bt[4] = e1.x, bt[5] = e1.y, bt[6] = e1.z, bt[7] = e1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) gm[j] = gamma[c0 + j], bt[j] = beta[c0 + j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cpg;
    const float mean = sstat[2 * g], rstd = sstat[2 * g + 1];
    a1[j] = rstd * gm[j];
    b1[j] = bt[j] - mean * rstd * gm[j];
    if (rs) {
      rs[j] = rstd;
      ms[j] = -mean * rstd;
    }
  }
}

// sum over the cluster of every CTA's part[0 .. 2G) -> tot[0 .. 2G) (tot: this CTA's scratch, >= (GN_CL + 1) * 2G floats).  The remote
// loads are spread over the threads (one dependent DSMEM round trip each instead of 2 * GN_CL in a row per thread).
__device__ __forceinline__ void gn_cluster_sum(const float* part, float* tot, int G) {
  const int n = 2 * G;
  float* tmp = tot + n;      // [GN_CL][2G]
  for (int i = threadIdx.x; i < n * GN_CL; i += blockDim.x) tmp[i] = gn_ld_peer(part + (i % n), (uint32_t)(i / n));
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < GN_CL; ++k) a += tmp[k * n + i];
    tot[i] = a;
  }
  __syncthreads();
}

// shared memory: [this CTA's rows, bf16][threads * 16 floats reduction scratch][G * 2 partials][G * 2 statistics]
struct GnClSmem {
  uint8_t* x;
  float* red;
  float* part;
  float* stat;
};
__device__ __forceinline__ GnClSmem gn_cl_carve(uint8_t* raw, int chunk_bytes, int threads, int G) {
  GnClSmem m;
  m.x = raw;
  m.red = reinterpret_cast<float*>(raw + chunk_bytes);
  m.part = m.red + threads * 16;
  m.stat = m.part + 2 * G;
  return m;
}
// fold per-thread per-channel pairs (s[j], q[j]) over the rpp row phases into per-channel pairs red[c * 2 + {0, 1}]
__device__ __forceinline__ void gn_fold_channels(float* red, const GnMap& m, int C, const float* s, const float* q) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[((m.rsub * C) + m.v * 8 + j) * 2] = s[j];
    red[((m.rsub * C) + m.v * 8 + j) * 2 + 1] = q[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += m.threads) {
    float a = red[c * 2], d = red[c * 2 + 1];
    for (int k = 1; k < m.rpp; ++k) {
      a += red[(k * C + c) * 2];
      d += red[(k * C + c) * 2 + 1];
    }
    red[c * 2] = a;
    red[c * 2 + 1] = d;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(320) gn_fwd_cluster_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             bf16* __restrict__ y, float* __restrict__ stats, int L, int C, int G,
                                                             int rows_per_cta, int chunk_bytes, float eps, int act, int vec) {
  extern __shared__ __align__(128) uint8_t gn_raw[];
  const GnMap m = gn_map(C);
  const GnClSmem sm = gn_cl_carve(gn_raw, chunk_bytes, m.threads, G);
  const int rank = (int)gn_cluster_rank();
  const int b = blockIdx.y;
  const int r0 = rank * rows_per_cta;
  const int nrows = max(0, min(L, r0 + rows_per_cta) - r0);
  const long long gbase = ((long long)b * L + r0) * C;
  // pass 1: stream this CTA's rows from HBM in batches of U per thread (all loads of a batch in flight at once), keep the raw
  // vectors in shared memory, accumulate the per-channel sums on the way.  (A bulk copy of the whole slice followed by a pass over
  // shared memory was measured first: one CTA per SM then has far too few bytes in flight -- 6 GB/s per SM.)
  bf16* sx = reinterpret_cast<bf16*>(sm.x);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  {
    constexpr int U = 8;
    const bf16* xg = x + gbase + m.v * 8;
    for (int r = m.rsub; r < nrows; r += U * m.rpp) {
      bf16x8 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * m.rpp < nrows) raw[u] = load8raw(xg + (size_t)(r + u * m.rpp) * C);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r + u * m.rpp < nrows) {
          st16(sx + (size_t)(r + u * m.rpp) * C + m.v * 8, raw[u]);
          float f[8];
          unpack8(raw[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += f[j];
            q[j] = fmaf(f[j], f[j], q[j]);
          }
        }
    }
  }
  gn_fold_channels(sm.red, m, C, s, q);
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += m.threads) {
    float a = 0.f, d = 0.f;
    for (int k = 0; k < cpg; ++k) {
      a += sm.red[(g * cpg + k) * 2];
      d += sm.red[(g * cpg + k) * 2 + 1];
    }
    sm.part[2 * g] = a;
    sm.part[2 * g + 1] = d;
  }
  gn_cluster_sync();      // every CTA's partials are written
  const float inv_count = 1.f / ((float)cpg * (float)L);
  gn_cluster_sum(sm.part, sm.red, G);      // red is free again: [2G] totals + [GN_CL][2G] staging
  for (int g = threadIdx.x; g < G; g += m.threads) {
    const float a = sm.red[2 * g], d = sm.red[2 * g + 1];
    const float mean = a * inv_count;
    const float var = fmaxf(d * inv_count - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    sm.stat[2 * g] = mean;
    sm.stat[2 * g + 1] = rstd;
    if (rank == 0) {
      stats[((long long)b * G + g) * 2] = mean;
      stats[((long long)b * G + g) * 2 + 1] = rstd;
    }
  }
  gn_cluster_sync();      // peers have read this CTA's partials (it may exit); the statistics are visible to the whole CTA
  float a1[8], b1[8];
  gn_coeffs_sm(sm.stat, gamma, beta, m.v * 8, cpg, vec != 0, a1, b1, nullptr, nullptr);
  bf16* yb = y + gbase + m.v * 8;
#pragma unroll 2
  for (int r = m.rsub; r < nrows; r += m.rpp) {
    float f[8];
    load8(sx + (size_t)r * C + m.v * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(f[j], a1[j], b1[j]);
      f[j] = act ? z * sigmoid_fast(z) : z;
    }
    store8(yb + (size_t)r * C, f);
  }
}

// rows per CTA, bytes of its resident slice and the dynamic shared memory of the cluster kernel; false if the sample does not fit
bool gn_cluster_fit(int L, int C, int G, int threads, int* rows_per_cta, int* chunk_bytes, size_t* smem) {
  static const bool off = getenv("PT_GN_NO_CLUSTER") != nullptr;
  if (off) return false;
  const int rows = (L + GN_CL - 1) / GN_CL;
  const int cb = (rows * C * 2 + 127) / 128 * 128;
  const size_t need = (size_t)cb + (size_t)threads * 16 * sizeof(float) + (size_t)4 * G * sizeof(float) + 16;
  *rows_per_cta = rows;
  *chunk_bytes = cb;
  *smem = need;
  return need <= 100 * 1024 && (C * 2) % 16 == 0;      // two CTAs per SM: all clusters of a batch of 32 resident at once (see above)
}

template <typename K, typename... Args>
int gn_cluster_launch(K kernel, int B, int threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(GN_CL, B, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GN_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (getenv("PT_GN_DEBUG")) {
    int n = 0;
    cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    fprintf(stderr, "[groupnorm] cluster kernel: %d threads, %zu B smem, %d clusters of %d resident at once (B = %d)\n", threads, smem, n, GN_CL, B);
  }
  PT_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, args...));
  PT_LAUNCH_CHECK();
  return PT_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C <= 2048).
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;

template <int NV, int ROWS>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              bf16* __restrict__ y, float* __restrict__ rowstats, long long M, int C, float eps) {
  // one warp per ROWS consecutive rows, all of them requested at once as raw bf16 (4 registers per 16-byte vector, converted
  // each time they are used): many rows in flight per warp at a register cost that still leaves three CTAs per SM
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  bf16x8 raw[ROWS][NV];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec && row0 + r < M) raw[r][i] = load8raw(x + (row0 + r) * C + v * 8);
    }
  }
  const float inv_c = 1.f / (float)C;
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float s = 0.f;
    if (row0 + r < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
          float f[8];
          unpack8(raw[r][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) s += f[j];
        }
      }
    }
    mean[r] = s;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) mean[r] = warp_sum(mean[r]) * inv_c;   // ROWS independent shuffle chains
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float q = 0.f;
    if (row0 + r < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
          float f[8];
          unpack8_again(raw[r][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - mean[r];
            q = fmaf(d, d, q);
          }
        }
      }
    }
    rstd[r] = q;
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) rstd[r] = rsqrtf(warp_sum(rstd[r]) * inv_c + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (row0 + r < M) {
          float f[8], o[8];
          unpack8_again(raw[r][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf((f[j] - mean[r]) * rstd[r], gm[j], bt[j]);
          store8(y + (row0 + r) * C + v * 8, o);
        }
      }
    }
  }
  if (lane < ROWS && row0 + lane < M) {
    float mu = mean[0], rs = rstd[0];
#pragma unroll
    for (int r = 1; r < ROWS; ++r)
      if (lane == r) mu = mean[r], rs = rstd[r];
    *reinterpret_cast<float2*>(rowstats + 2 * (row0 + lane)) = make_float2(mu, rs);
  }
}

// Each warp walks rows w, w + nwarps_total, ...; per-lane dgamma/dbeta partials stay in registers and are flushed once through
// shared memory + global atomics.  A row is held as raw bf16 (x, dy, dx_add: 12 registers per 8 channels) and the NEXT row of the
// warp is requested before the current one is reduced, so the two warp-wide reductions and the arithmetic of a row overlap the
// memory round trip of the following one.
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ rowstats,
                              const float* __restrict__ gamma, const bf16* __restrict__ dx_add, bf16* __restrict__ dx,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C) {
  extern __shared__ float sh[];  // [warps][2*C]: one private slot per warp, no atomics
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)C;
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[i][j] = ab[i][j] = 0.f;
  }
  bf16x8 cx[NV], cd[NV], ca[NV];
  float2 cst = make_float2(0.f, 0.f);
  auto fetch = [&](long long row, bf16x8* px, bf16x8* pd, bf16x8* pa, float2& st) {
    if (row < M) {
      st = __ldg(reinterpret_cast<const float2*>(rowstats + 2 * row));
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
          px[i] = load8raw(x + row * C + v * 8);
          pd[i] = load8raw(dy + row * C + v * 8);
          if (dx_add) pa[i] = load8raw(dx_add + row * C + v * 8);
        }
      }
    }
  };
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  fetch(row, cx, cd, ca, cst);
  for (; row < M; row += wstride) {
    bf16x8 nx[NV], nd[NV], na[NV];
    float2 nst = make_float2(0.f, 0.f);
    fetch(row + wstride, nx, nd, na, nst);
    const float mean = cst.x, rstd = cst.y;
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float fx[8], fd[8];
        unpack8(cx[i], fx);
        unpack8(cd[i], fd);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (fx[j] - mean) * rstd;
          ag[i][j] = fmaf(fd[j], xh, ag[i][j]);
          ab[i][j] += fd[j];
          const float dg = fd[j] * gm[j];
          c1 += dg;
          c2 = fmaf(dg, xh, c2);
        }
      }
    }
    c1 = warp_sum(c1) * inv_c;
    c2 = warp_sum(c2) * inv_c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float fx[8], fd[8], o[8];
        unpack8_again(cx[i], fx);
        unpack8_again(cd[i], fd);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (fx[j] - mean) * rstd;
          o[j] = rstd * (fd[j] * gm[j] - c1 - xh * c2);
        }
        if (dx_add) {
          float t[8];
          unpack8(ca[i], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += t[j];
        }
        store8(dx + row * C + v * 8, o);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cx[i] = nx[i];
      cd[i] = nd[i];
      ca[i] = na[i];
    }
    cst = nst;
  }
  float* slot = sh + (threadIdx.x >> 5) * 2 * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        slot[v * 8 + j] = ag[i][j];
        slot[C + v * 8 + j] = ab[i][j];
      }
    }
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += sh[w * 2 * C + i];
    atomicAdd(i < C ? &dgamma[i] : &dbeta[i - C], a);
  }
}

}  // namespace

extern "C" int pt_groupnorm_stats(const void* x, float* stats, int B, int L, int C, int G, float eps, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_stats: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  PT_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * B * G, st));
  gn_stats_kernel<<<dim3(g.chunks_red, B), g.threads, g.threads * 16 * sizeof(float), st>>>((const bf16*)x, stats, L, C, G, g.rows_red);
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
  if (gn_fast(stats, gamma, beta, C, G))
    gn_apply_kernel<GN_U_APPLY, true><<<dim3(g.chunks_apply, B), g.threads, 0, (cudaStream_t)stream>>>((const bf16*)x, stats, gamma, beta,
                                                                                                      (bf16*)y, L, C, G, g.rows_apply, act);
  else
    gn_apply_kernel<GN_U_APPLY, false><<<dim3(g.chunks_apply, B), g.threads, 0, (cudaStream_t)stream>>>((const bf16*)x, stats, gamma, beta,
                                                                                                       (bf16*)y, L, C, G, g.rows_apply, act);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

// statistics + normalisation (+ SiLU) in one call: the sample-resident cluster kernel when a sample fits, else the two streaming kernels
extern "C" int pt_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int B, int L, int C, int G,
                                float eps, int act, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_fwd: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  int rows, cb;
  size_t smem;
  if (gn_cluster_fit(L, C, G, g.threads, &rows, &cb, &smem)) {
    PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(gn_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int vec = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    return gn_cluster_launch(gn_fwd_cluster_kernel, B, g.threads, smem, (cudaStream_t)stream, (const bf16*)x, gamma, beta, (bf16*)y, stats, L, C, G,
                             rows, cb, eps, act, vec);
  }
  if (int r = pt_groupnorm_stats(x, stats, B, L, C, G, eps, stream)) return r;
  return pt_groupnorm_apply(x, stats, gamma, beta, y, B, L, C, G, act, stream);
}

extern "C" int pt_groupnorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta, const void* dx_add,
                                void* dx, float* dgamma, float* dbeta, float* scratch, int B, int L, int C, int G, int act, void* stream) {
  PT_REQUIRE(B > 0 && L > 0 && G > 0 && C % G == 0, "groupnorm_bwd: B=%d L=%d C=%d G=%d", B, L, C, G);
  GnGeom g;
  if (int r = gn_geom(B, L, C, &g)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  const int fast = gn_fast(stats, gamma, beta, C, G);
  PT_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * B * G, st));
  const dim3 gr(g.chunks_red, B), ga(g.chunks_bapply, B);
  const size_t smem = g.threads * 16 * sizeof(float);
  const bf16 *dyp = (const bf16*)dy, *xp = (const bf16*)x, *addp = (const bf16*)dx_add;
  if (fast) {
    gn_bwd_reduce_kernel<true><<<gr, g.threads, smem, st>>>(dyp, xp, stats, gamma, beta, dgamma, dbeta, scratch, L, C, G, g.rows_red, act);
    PT_LAUNCH_CHECK();
    gn_bwd_apply_kernel<true><<<ga, g.threads, 0, st>>>(dyp, xp, stats, gamma, beta, scratch, addp, (bf16*)dx, L, C, G, g.rows_bapply, act);
  } else {
    gn_bwd_reduce_kernel<false><<<gr, g.threads, smem, st>>>(dyp, xp, stats, gamma, beta, dgamma, dbeta, scratch, L, C, G, g.rows_red, act);
    PT_LAUNCH_CHECK();
    gn_bwd_apply_kernel<false><<<ga, g.threads, 0, st>>>(dyp, xp, stats, gamma, beta, scratch, addp, (bf16*)dx, L, C, G, g.rows_bapply, act);
  }
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* rowstats, int64_t M, int C, float eps,
                                void* stream) {
  PT_REQUIRE(M > 0 && C % 8 == 0 && C / 8 <= 32 * LN_MAXV, "layernorm_fwd: M=%lld C=%d", (long long)M, C);
  PT_REQUIRE(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "layernorm_fwd: gamma/beta must be 16-byte aligned");
  const int wpb = 8;
  const int nv = (C / 8 + 31) / 32;
#define LN_FWD(NV_, R_)                                                                                                                   \
  case NV_:                                                                                                                               \
    ln_fwd_kernel<NV_, R_><<<(unsigned)((M + wpb * R_ - 1) / (wpb * R_)), wpb * 32, 0, (cudaStream_t)stream>>>((const bf16*)x, gamma, beta, \
                                                                                                              (bf16*)y, rowstats, M, C, eps); \
    break;
  switch (nv) {
    LN_FWD(1, 8) LN_FWD(2, 8) LN_FWD(3, 4) LN_FWD(4, 4) LN_FWD(5, 2) LN_FWD(6, 2) LN_FWD(7, 2) LN_FWD(8, 2)
  }
#undef LN_FWD
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_layernorm_bwd(const void* dy, const void* x, const float* rowstats, const float* gamma, const void* dx_add, void* dx,
                                float* dgamma, float* dbeta, int64_t M, int C, void* stream) {
  PT_REQUIRE(M > 0 && C % 8 == 0 && C / 8 <= 32 * LN_MAXV, "layernorm_bwd: M=%lld C=%d", (long long)M, C);
  PT_REQUIRE((reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(rowstats) & 7) == 0,
             "layernorm_bwd: gamma must be 16-byte aligned, rowstats 8-byte aligned");
  const int wpb = 8;
  long long blocks = (M + wpb - 1) / wpb;
  const int nv = (C / 8 + 31) / 32;
  // every CTA ends with a 2C-value flush (shared-memory fold + global atomics), so the grid is exactly one resident wave: as
  // many CTAs as fit at once (register-limited: two per SM up to C = 512, one above), each warp walking many rows
  const size_t smem = (size_t)wpb * 2 * C * sizeof(float);
#define LN_BWD(NV_)                                                                                                            \
  case NV_: {                                                                                                                  \
    static int occ = 0;                                                                                                        \
    if (!occ) {                                                                                                                \
      PT_CUDA_OK(cudaFuncSetAttribute(ln_bwd_kernel<NV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * NV_ * 256 * 4)); \
      int o = 0;                                                                                                               \
      PT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, ln_bwd_kernel<NV_>, wpb * 32, 8 * 2 * NV_ * 256 * 4));       \
      occ = o < 1 ? 1 : (o > 2 ? 2 : o);                                                                                       \
    }                                                                                                                          \
    const long long cap = (long long)occ * pt_num_sms();                                                                       \
    if (blocks > cap) blocks = cap;                                                                                            \
    ln_bwd_kernel<NV_><<<(unsigned)blocks, wpb * 32, smem, (cudaStream_t)stream>>>(                                             \
        (const bf16*)dy, (const bf16*)x, rowstats, gamma, (const bf16*)dx_add, (bf16*)dx, dgamma, dbeta, M, C);                \
  } break;
  switch (nv) {
    LN_BWD(1) LN_BWD(2) LN_BWD(3) LN_BWD(4) LN_BWD(5) LN_BWD(6) LN_BWD(7) LN_BWD(8)
  }
#undef LN_BWD
  PT_LAUNCH_CHECK();
  return PT_OK;
}
