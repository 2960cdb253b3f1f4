// seanet_core.h -- the per-thread bodies and launch geometry of the SEANet kernels (libpt_seanet.so).
//
// Every kernel of this library but one is "one thread = one register tile of outputs", with no shared memory and no barriers, so
// a kernel is completely described by  body(params, blockIdx, threadIdx.x, blockDim.x)  plus the grid that covers the problem.
// Both live here as plain inline functions.  (The exception, the whole-sequence LSTM kernel at the end of this file, is three such
// bodies separated by barriers.)  nvcc compiles them as device code (seanet.cu wraps each in a __global__ that
// passes the built-in indices); tests/seanet_emul.cpp compiles THE SAME text with g++ and walks the grid in a host loop, which
// lets the CPU-only test tier check the index arithmetic (padding, strides, phase decomposition, packing) against the oracle.
// That host build is a checker of this file, not a code path of the product: nothing under prompt_tts_b200/ loads it.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "../../../include/prompt_tts_seanet.h"

#ifdef __CUDACC__
#define SN_HD __host__ __device__ __forceinline__
#else
#define SN_HD inline
#endif

struct sn_f4 {
  float v[4];
};
struct sn_grid {
  unsigned x, y, z;
};

#if defined(__CUDA_ARCH__)
#define SN_LD(p) __ldg(p)
SN_HD sn_f4 sn_ld4(const float* p) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  sn_f4 r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
#else
#define SN_LD(p) (*(p))
SN_HD sn_f4 sn_ld4(const float* p) {
  sn_f4 r;
  r.v[0] = p[0]; r.v[1] = p[1]; r.v[2] = p[2]; r.v[3] = p[3];
  return r;
}
#endif

// 16-byte loads / stores of ordinary (here: shared) memory
SN_HD sn_f4 sn_ldv4(const float* p) {
  sn_f4 r;
#if defined(__CUDA_ARCH__)
  const float4 t = *reinterpret_cast<const float4*>(p);
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
#else
  r.v[0] = p[0]; r.v[1] = p[1]; r.v[2] = p[2]; r.v[3] = p[3];
#endif
  return r;
}
SN_HD void sn_stv4(float* p, const sn_f4& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float4*>(p) = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
#else
  p[0] = v.v[0]; p[1] = v.v[1]; p[2] = v.v[2]; p[3] = v.v[3];
#endif
}

SN_HD float sn_elu(float v) { return v > 0.f ? v : expm1f(v); }
SN_HD float sn_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }
SN_HD unsigned sn_cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

constexpr int SN_THREADS = 128;  // threads per block of every kernel
constexpr int SN_CT = 8;         // conv: output channels per thread
constexpr int SN_TT = 4;         // conv: output samples per thread (blockDim.x apart: a warp reads consecutive samples)
constexpr int SN_UC = 4;         // transposed conv: output channels per thread
constexpr int SN_UR = 2;         // transposed conv: output phases (t mod stride) per thread
constexpr int SN_UT = 4;         // transposed conv: input frames per thread
constexpr int SN_LR = 4;         // linear_rows: rows per thread (x 4 columns)
constexpr int SN_SR = 2;         // lstm_step: sequences per thread (x 1 hidden unit x 4 gates)

// ------------------------------------------------------------------------------------------------ Conv1d
// Index into the unpadded signal for position q (already shifted by -pad_left); -1 = a zero sample.
SN_HD int sn_src_index(int q, int L, int reflect) {
  if (q < 0) return reflect ? -q : -1;
  if (q >= L) return reflect ? 2 * (L - 1) - q : -1;
  return q;
}

SN_HD sn_grid sn_conv1d_grid(const pt_sn_conv_t& p) {
  sn_grid g;
  g.x = sn_cdiv(p.Lout, SN_THREADS * SN_TT);
  g.y = sn_cdiv(p.Co, SN_CT);
  g.z = (unsigned)p.B;
  return g;
}

// thread (bx, tx) -> samples t_j = bx * ntx * SN_TT + tx + j * ntx; by -> channels [by * SN_CT, +SN_CT); bz -> batch entry.
// The weights of a (ci, k) step are the same for the whole warp (broadcast loads); x is read along t.
SN_HD void sn_conv1d_thread(const pt_sn_conv_t& p, int bx, int by, int bz, int tx, int ntx) {
  const int co0 = by * SN_CT;
  const int t0 = bx * ntx * SN_TT + tx;
  float acc[SN_CT][SN_TT];
#pragma unroll
  for (int c = 0; c < SN_CT; ++c)
#pragma unroll
    for (int j = 0; j < SN_TT; ++j) acc[c][j] = 0.f;
  int q0[SN_TT];  // t * stride - pad_left, or a value that keeps every tap out of range for t >= Lout
#pragma unroll
  for (int j = 0; j < SN_TT; ++j) q0[j] = (t0 + j * ntx) * p.stride - p.pad_left;
  const float* xb = p.x + (size_t)bz * p.Ci * p.Lin;
  const size_t wrow = (size_t)p.Ci * p.K;
  for (int ci = 0; ci < p.Ci; ++ci) {
    const float* xc = xb + (size_t)ci * p.Lin;
    const float* wc = p.w + (size_t)co0 * wrow + (size_t)ci * p.K;
    for (int k = 0; k < p.K; ++k) {
      float wv[SN_CT];
#pragma unroll
      for (int c = 0; c < SN_CT; ++c) wv[c] = (co0 + c < p.Co) ? SN_LD(wc + (size_t)c * wrow + k) : 0.f;
#pragma unroll
      for (int j = 0; j < SN_TT; ++j) {
        float xv = 0.f;
        if (t0 + j * ntx < p.Lout) {
          const int s = sn_src_index(q0[j] + k * p.dil, p.Lin, p.reflect);
          if (s >= 0) xv = SN_LD(xc + s);
        }
#pragma unroll
        for (int c = 0; c < SN_CT; ++c) acc[c][j] = fmaf(wv[c], xv, acc[c][j]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SN_CT; ++c) {
    const int co = co0 + c;
    if (co >= p.Co) continue;
    const float bv = p.bias ? SN_LD(p.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < SN_TT; ++j) {
      const int t = t0 + j * ntx;
      if (t >= p.Lout) continue;
      const size_t o = ((size_t)bz * p.Co + co) * p.Lout + t;
      float v = acc[c][j] + bv;
      if (p.res) v += SN_LD(p.res + o);
      if (p.y) p.y[o] = v;
      if (p.y_elu) p.y_elu[o] = sn_elu(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ ConvTranspose1d
// Output sample tf of the FULL transposed convolution decomposes as tf = i * stride + r (frame i, phase r); it receives
//   sum_ci sum_m x[ci, i - m] * w[ci, co, r + m * stride]     for the taps r + m * stride < K  (two taps when K = 2 * stride).
// A thread owns SN_UT frames x SN_UR phases x SN_UC channels: the weights it needs are warp-uniform, x is read along i.
SN_HD int sn_convtr_phase_groups(const pt_sn_conv_t& p) { return (p.stride + SN_UR - 1) / SN_UR; }
SN_HD int sn_convtr_frames(const pt_sn_conv_t& p) { return (p.pad_left + p.Lout + p.stride - 1) / p.stride; }  // i in [0, this)

SN_HD sn_grid sn_convtr_grid(const pt_sn_conv_t& p) {
  sn_grid g;
  g.x = sn_cdiv(sn_convtr_frames(p), SN_THREADS * SN_UT);
  g.y = (unsigned)sn_convtr_phase_groups(p) * sn_cdiv(p.Co, SN_UC);
  g.z = (unsigned)p.B;
  return g;
}

SN_HD void sn_convtr_thread(const pt_sn_conv_t& p, int bx, int by, int bz, int tx, int ntx) {
  const int npg = sn_convtr_phase_groups(p);
  const int r0 = (by % npg) * SN_UR;
  const int co0 = (by / npg) * SN_UC;
  const int i0 = bx * ntx * SN_UT + tx;
  const int taps = (p.K + p.stride - 1) / p.stride;
  float acc[SN_UT][SN_UR][SN_UC];
#pragma unroll
  for (int j = 0; j < SN_UT; ++j)
#pragma unroll
    for (int r = 0; r < SN_UR; ++r)
#pragma unroll
      for (int c = 0; c < SN_UC; ++c) acc[j][r][c] = 0.f;
  const float* xb = p.x + (size_t)bz * p.Ci * p.Lin;
  for (int ci = 0; ci < p.Ci; ++ci) {
    const float* xc = xb + (size_t)ci * p.Lin;
    const float* wc = p.w + ((size_t)ci * p.Co + co0) * p.K;
    for (int m = 0; m < taps; ++m) {
      float xv[SN_UT];
#pragma unroll
      for (int j = 0; j < SN_UT; ++j) {
        const int s = i0 + j * ntx - m;
        xv[j] = (s >= 0 && s < p.Lin) ? SN_LD(xc + s) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < SN_UR; ++r) {
        const int k = r0 + r + m * p.stride;
        if (r0 + r >= p.stride || k >= p.K) continue;
#pragma unroll
        for (int c = 0; c < SN_UC; ++c) {
          const float wv = (co0 + c < p.Co) ? SN_LD(wc + (size_t)c * p.K + k) : 0.f;
#pragma unroll
          for (int j = 0; j < SN_UT; ++j) acc[j][r][c] = fmaf(xv[j], wv, acc[j][r][c]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SN_UC; ++c) {
    const int co = co0 + c;
    if (co >= p.Co) continue;
    const float bv = p.bias ? SN_LD(p.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < SN_UT; ++j)
#pragma unroll
      for (int r = 0; r < SN_UR; ++r) {
        if (r0 + r >= p.stride) continue;
        const long long t = (long long)(i0 + j * ntx) * p.stride + r0 + r - p.pad_left;
        if (t < 0 || t >= p.Lout) continue;
        const size_t o = ((size_t)bz * p.Co + co) * p.Lout + (size_t)t;
        float v = acc[j][r][c] + bv;
        if (p.res) v += SN_LD(p.res + o);
        if (p.y) p.y[o] = v;
        if (p.y_elu) p.y_elu[o] = sn_elu(v);
      }
  }
}

// ------------------------------------------------------------------------------------------------ element-indexed kernels
SN_HD sn_grid sn_linear_grid_1d(long long n) {
  sn_grid g;
  g.x = sn_cdiv(n, SN_THREADS);
  g.y = g.z = 1;
  return g;
}

// w[r, :] = g[r] * v[r, :] / |v[r, :]| -- one thread per row (load-time work; the sum of squares is kept in double)
SN_HD void sn_weight_norm_thread(const float* v, const float* g, float* w, int rows, int cols, int bx, int tx, int ntx) {
  const int r = bx * ntx + tx;
  if (r >= rows) return;
  const float* vr = v + (size_t)r * cols;
  double ss = 0.0;
  for (int i = 0; i < cols; ++i) {
    const double a = (double)SN_LD(vr + i);
    ss += a * a;
  }
  const float scale = (float)((double)SN_LD(g + r) / sqrt(ss));
  for (int i = 0; i < cols; ++i) w[(size_t)r * cols + i] = SN_LD(vr + i) * scale;
}

// wt4[k][j][q] = W[q * H + j][k]
SN_HD void sn_lstm_pack_thread(const float* w, float* wt4, int H, int bx, int tx, int ntx) {
  const long long o = (long long)bx * ntx + tx;
  if (o >= 4LL * H * H) return;
  const int q = (int)(o & 3);
  const int j = (int)((o >> 2) % H);
  const int k = (int)((o >> 2) / H);
  wt4[o] = SN_LD(w + ((size_t)q * H + j) * H + k);
}

SN_HD void sn_lstm_pack_bias_thread(const float* b_ih, const float* b_hh, float* bias4, int H, int bx, int tx, int ntx) {
  const int o = bx * ntx + tx;
  if (o >= 4 * H) return;
  const int q = o & 3, j = o >> 2;
  bias4[o] = SN_LD(b_ih + q * H + j) + SN_LD(b_hh + q * H + j);
}

// out[t, b, c] = x[b, c, t]
SN_HD void sn_ncl_to_tbc_thread(const float* x, float* out, int B, int Cn, int T, int bx, int tx, int ntx) {
  const long long o = (long long)bx * ntx + tx;
  if (o >= (long long)B * Cn * T) return;
  const int c = (int)(o % Cn);
  const int b = (int)((o / Cn) % B);
  const int t = (int)(o / ((long long)Cn * B));
  out[o] = SN_LD(x + ((size_t)b * Cn + c) * T + t);
}

// y[b, c, t] = hseq[t, b, c] + x[b, c, t]
SN_HD void sn_tbc_add_to_ncl_thread(const float* hseq, const float* x, float* y, float* y_elu, int B, int Cn, int T, int bx, int tx,
                                    int ntx) {
  const long long o = (long long)bx * ntx + tx;
  if (o >= (long long)B * Cn * T) return;
  const int t = (int)(o % T);
  const int c = (int)((o / T) % Cn);
  const int b = (int)(o / ((long long)T * Cn));
  const float v = SN_LD(hseq + ((size_t)t * B + b) * Cn + c) + SN_LD(x + o);
  if (y) y[o] = v;
  if (y_elu) y_elu[o] = sn_elu(v);
}

// ------------------------------------------------------------------------------------------------ row-major GEMM for the LSTM input projection
// out[r, n] = bias[n] + sum_k a[r, k] * wt[k, n]; a thread owns SN_LR rows x 4 consecutive columns; lanes run along n, so the
// weight loads are coalesced 16-byte loads and the a loads are warp-uniform.  Kd % 4 == 0, N % 4 == 0.
SN_HD sn_grid sn_linear_rows_grid(int R, int N) {
  sn_grid g;
  g.x = sn_cdiv(N / 4, SN_THREADS);
  g.y = sn_cdiv(R, SN_LR);
  g.z = 1;
  return g;
}

SN_HD void sn_linear_rows_thread(const float* a, const float* wt, const float* bias, float* out, int R, int Kd, int N, int bx, int by,
                                 int tx, int ntx) {
  const int n4 = bx * ntx + tx;
  if (n4 * 4 >= N) return;
  const int r0 = by * SN_LR;
  float acc[SN_LR][4];
  const sn_f4 bv = bias ? sn_ld4(bias + n4 * 4) : sn_f4{{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int r = 0; r < SN_LR; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = bv.v[q];
  const float* ar[SN_LR];
#pragma unroll
  for (int r = 0; r < SN_LR; ++r) ar[r] = a + (size_t)(r0 + r < R ? r0 + r : R - 1) * Kd;
  for (int k = 0; k < Kd; k += 4) {
    sn_f4 av[SN_LR];
#pragma unroll
    for (int r = 0; r < SN_LR; ++r) av[r] = sn_ld4(ar[r] + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const sn_f4 wv = sn_ld4(wt + (size_t)(k + kk) * N + n4 * 4);
#pragma unroll
      for (int r = 0; r < SN_LR; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(av[r].v[kk], wv.v[q], acc[r][q]);
    }
  }
#pragma unroll
  for (int r = 0; r < SN_LR; ++r) {
    if (r0 + r >= R) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) out[(size_t)(r0 + r) * N + n4 * 4 + q] = acc[r][q];
  }
}

// ------------------------------------------------------------------------------------------------ LSTM time step
// thread = hidden unit j of SN_SR sequences: its four gate pre-activations are one 16-byte row of the packed weights per k.
SN_HD sn_grid sn_lstm_step_grid(int B, int H) {
  sn_grid g;
  g.x = sn_cdiv(H, SN_THREADS);
  g.y = sn_cdiv(B, SN_SR);
  g.z = 1;
  return g;
}

SN_HD void sn_lstm_step_thread(const float* xg, const float* whh_t4, float* hseq, float* c, int t, int B, int H, int bx, int by, int tx,
                               int ntx) {
  const int j = bx * ntx + tx;
  if (j >= H) return;
  const int b0 = by * SN_SR;
  float acc[SN_SR][4];
  int bb[SN_SR];
#pragma unroll
  for (int r = 0; r < SN_SR; ++r) {
    bb[r] = b0 + r < B ? b0 + r : B - 1;
    const sn_f4 g = sn_ld4(xg + (((size_t)t * B + bb[r]) * H + j) * 4);
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = g.v[q];
  }
  if (t > 0) {
    const float* hp = hseq + (size_t)(t - 1) * B * H;
    for (int k = 0; k < H; k += 4) {
      sn_f4 hv[SN_SR];
#pragma unroll
      for (int r = 0; r < SN_SR; ++r) hv[r] = sn_ld4(hp + (size_t)bb[r] * H + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const sn_f4 wv = sn_ld4(whh_t4 + ((size_t)(k + kk) * H + j) * 4);
#pragma unroll
        for (int r = 0; r < SN_SR; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(hv[r].v[kk], wv.v[q], acc[r][q]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < SN_SR; ++r) {
    if (b0 + r >= B) continue;
    const size_t o = (size_t)bb[r] * H + j;
    const float cp = t > 0 ? c[o] : 0.f;
    const float cn = sn_sigmoid(acc[r][1]) * cp + sn_sigmoid(acc[r][0]) * tanhf(acc[r][2]);
    c[o] = cn;
    hseq[(size_t)t * B * H + o] = sn_sigmoid(acc[r][3]) * tanhf(cn);
  }
}

// ================================================================================================ packed-weight kernels
// wp[(ci * K + k) * Cop + co]: a thread's output channels are consecutive floats, read as 16-byte loads that are the same for the
// whole warp; rows are zero-filled up to Cop, so the inner loops carry no channel predicate.
SN_HD void sn_pack_conv_weight_thread(const float* w, float* wp, int Co, int Ci, int K, int Cop, int transposed, int bx, int tx,
                                      int ntx) {
  const long long o = (long long)bx * ntx + tx;
  if (o >= (long long)Ci * K * Cop) return;
  const int co = (int)(o % Cop);
  const int k = (int)((o / Cop) % K);
  const int ci = (int)(o / ((long long)Cop * K));
  float v = 0.f;
  if (co < Co) v = transposed ? SN_LD(w + ((size_t)ci * Co + co) * K + k) : SN_LD(w + ((size_t)co * Ci + ci) * K + k);
  wp[o] = v;
}

SN_HD int sn_conv1d_packed_ct(const pt_sn_conv_t& p) { return p.Co_pad % 16 == 0 ? 16 : 8; }  // output channels per thread

template <int CT>
SN_HD sn_grid sn_conv1d_packed_grid(const pt_sn_conv_t& p) {
  sn_grid g;
  g.x = sn_cdiv(p.Lout, SN_THREADS * SN_TT);
  g.y = (unsigned)(p.Co_pad / CT);
  g.z = (unsigned)p.B;
  return g;
}

template <int CT>
SN_HD void sn_conv1d_packed_thread(const pt_sn_conv_t& p, int bx, int by, int bz, int tx, int ntx) {
  const int co0 = by * CT;
  const int t0 = bx * ntx * SN_TT + tx;
  const int Cop = p.Co_pad;
  float acc[CT][SN_TT];
#pragma unroll
  for (int c = 0; c < CT; ++c)
#pragma unroll
    for (int j = 0; j < SN_TT; ++j) acc[c][j] = 0.f;
  int q0[SN_TT];
  bool interior = true;  // every tap of every sample of this thread lies inside the signal
#pragma unroll
  for (int j = 0; j < SN_TT; ++j) {
    const int t = t0 + j * ntx;
    q0[j] = t * p.stride - p.pad_left;
    interior = interior && t < p.Lout && q0[j] >= 0 && q0[j] + (p.K - 1) * p.dil < p.Lin;
  }
  const float* xb = p.x + (size_t)bz * p.Ci * p.Lin;
  const float* wb = p.w + co0;
  if (interior) {
    for (int ci = 0; ci < p.Ci; ++ci) {
      const float* xc = xb + (size_t)ci * p.Lin;
      const float* wc = wb + (size_t)ci * p.K * Cop;
      for (int k = 0; k < p.K; ++k) {
        float wv[CT];
#pragma unroll
        for (int c4 = 0; c4 < CT / 4; ++c4) {
          const sn_f4 w4 = sn_ld4(wc + (size_t)k * Cop + c4 * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) wv[c4 * 4 + q] = w4.v[q];
        }
        const int off = k * p.dil;
#pragma unroll
        for (int j = 0; j < SN_TT; ++j) {
          const float xv = SN_LD(xc + q0[j] + off);
#pragma unroll
          for (int c = 0; c < CT; ++c) acc[c][j] = fmaf(wv[c], xv, acc[c][j]);
        }
      }
    }
  } else {
    const int last = p.Lin - 1;
    for (int ci = 0; ci < p.Ci; ++ci) {
      const float* xc = xb + (size_t)ci * p.Lin;
      const float* wc = wb + (size_t)ci * p.K * Cop;
      for (int k = 0; k < p.K; ++k) {
        float wv[CT];
#pragma unroll
        for (int c4 = 0; c4 < CT / 4; ++c4) {
          const sn_f4 w4 = sn_ld4(wc + (size_t)k * Cop + c4 * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) wv[c4 * 4 + q] = w4.v[q];
        }
#pragma unroll
        for (int j = 0; j < SN_TT; ++j) {
          // branch-free form of sn_src_index: mirror, clamp so the load is always legal, then select
          const int q = q0[j] + k * p.dil;
          const bool inside = q >= 0 && q <= last;
          int s = q < 0 ? -q : q;
          s = s > last ? 2 * last - s : s;
          s = s < 0 ? 0 : (s > last ? last : s);
          const bool use = (t0 + j * ntx < p.Lout) && (inside || p.reflect != 0);
          const float xl = SN_LD(xc + s);
          const float xv = use ? xl : 0.f;
#pragma unroll
          for (int c = 0; c < CT; ++c) acc[c][j] = fmaf(wv[c], xv, acc[c][j]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    const int co = co0 + c;
    if (co >= p.Co) continue;
    const float bv = p.bias ? SN_LD(p.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < SN_TT; ++j) {
      const int t = t0 + j * ntx;
      if (t >= p.Lout) continue;
      const size_t o = ((size_t)bz * p.Co + co) * p.Lout + t;
      float v = acc[c][j] + bv;
      if (p.res) v += SN_LD(p.res + o);
      if (p.y) p.y[o] = v;
      if (p.y_elu) p.y_elu[o] = sn_elu(v);
    }
  }
}

// transposed convolution, packed weights: SN_UT frames x SN_UR phases x SN_PC channels per thread
constexpr int SN_PC = 8;

SN_HD sn_grid sn_convtr_packed_grid(const pt_sn_conv_t& p) {
  sn_grid g;
  g.x = sn_cdiv(sn_convtr_frames(p), SN_THREADS * SN_UT);
  g.y = (unsigned)sn_convtr_phase_groups(p) * (unsigned)(p.Co_pad / SN_PC);
  g.z = (unsigned)p.B;
  return g;
}

SN_HD void sn_convtr_packed_thread(const pt_sn_conv_t& p, int bx, int by, int bz, int tx, int ntx) {
  const int npg = sn_convtr_phase_groups(p);
  const int r0 = (by % npg) * SN_UR;
  const int co0 = (by / npg) * SN_PC;
  const int i0 = bx * ntx * SN_UT + tx;
  const int taps = (p.K + p.stride - 1) / p.stride;
  const int Cop = p.Co_pad;
  const int last = p.Lin - 1;
  float acc[SN_UT][SN_UR][SN_PC];
#pragma unroll
  for (int j = 0; j < SN_UT; ++j)
#pragma unroll
    for (int r = 0; r < SN_UR; ++r)
#pragma unroll
      for (int c = 0; c < SN_PC; ++c) acc[j][r][c] = 0.f;
  const float* xb = p.x + (size_t)bz * p.Ci * p.Lin;
  for (int ci = 0; ci < p.Ci; ++ci) {
    const float* xc = xb + (size_t)ci * p.Lin;
    const float* wc = p.w + (size_t)ci * p.K * Cop + co0;
    for (int m = 0; m < taps; ++m) {
      float xv[SN_UT];
#pragma unroll
      for (int j = 0; j < SN_UT; ++j) {
        const int s = i0 + j * ntx - m;
        const int sc = s < 0 ? 0 : (s > last ? last : s);
        const float xl = SN_LD(xc + sc);
        xv[j] = (s >= 0 && s <= last) ? xl : 0.f;
      }
#pragma unroll
      for (int r = 0; r < SN_UR; ++r) {
        const int k = r0 + r + m * p.stride;
        if (r0 + r >= p.stride || k >= p.K) continue;  // warp-uniform
#pragma unroll
        for (int c4 = 0; c4 < SN_PC / 4; ++c4) {
          const sn_f4 w4 = sn_ld4(wc + (size_t)k * Cop + c4 * 4);
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < SN_UT; ++j) acc[j][r][c4 * 4 + q] = fmaf(xv[j], w4.v[q], acc[j][r][c4 * 4 + q]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SN_PC; ++c) {
    const int co = co0 + c;
    if (co >= p.Co) continue;
    const float bv = p.bias ? SN_LD(p.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < SN_UT; ++j)
#pragma unroll
      for (int r = 0; r < SN_UR; ++r) {
        if (r0 + r >= p.stride) continue;
        const long long t = (long long)(i0 + j * ntx) * p.stride + r0 + r - p.pad_left;
        if (t < 0 || t >= p.Lout) continue;
        const size_t o = ((size_t)bz * p.Co + co) * p.Lout + (size_t)t;
        float v = acc[j][r][c] + bv;
        if (p.res) v += SN_LD(p.res + o);
        if (p.y) p.y[o] = v;
        if (p.y_elu) p.y_elu[o] = sn_elu(v);
      }
  }
}

// ================================================================================================ LSTM, whole sequence in one launch
// Block bx owns SN_PU hidden units for all sequences and all steps: warp u = unit bx * SN_PU + u, lane = sequence within a chunk
// of 32.  Per block in shared memory: the units' recurrent weights wsm[u][k][gate] (loaded once) and the previous hidden state of
// the current chunk hs[32][H + 4] (rows padded by 4 floats: the 16-byte reads of a quarter-warp fall into distinct banks).
// The three pieces below are separated by barriers in the kernel (block barrier after load / stage / compute, grid barrier per
// step); the host-side checker calls them in the same order.
constexpr int SN_PU = 4;
#if defined(__CUDA_ARCH__)
#define SN_LD4_CG(p, r)                                            \
  {                                                                \
    const float4 t_ = __ldcg(reinterpret_cast<const float4*>(p)); \
    (r).v[0] = t_.x; (r).v[1] = t_.y; (r).v[2] = t_.z; (r).v[3] = t_.w; \
  }
#else
#define SN_LD4_CG(p, r) \
  { (r).v[0] = (p)[0]; (r).v[1] = (p)[1]; (r).v[2] = (p)[2]; (r).v[3] = (p)[3]; }
#endif

SN_HD size_t sn_lstm_seq_smem_floats(int H) { return (size_t)SN_PU * H * 4 + (size_t)32 * (H + 4); }

SN_HD void sn_lstm_seq_load_w(const float* whh_t4, float* wsm, int H, int bx, int tx, int ntx) {
  for (int i = tx; i < SN_PU * H; i += ntx) {
    const int u = i / H, k = i % H;
    sn_stv4(wsm + (size_t)i * 4, sn_ld4(whh_t4 + ((size_t)k * H + (size_t)bx * SN_PU + u) * 4));
  }
}

// hs[r][:] = hseq[t - 1][b0 + r][:] for r < nb, read through L2 (written by other blocks in the previous step)
SN_HD void sn_lstm_seq_stage(const float* hseq, float* hs, int t, int b0, int nb, int B, int H, int tx, int ntx) {
  const int h4 = H / 4;
  const float* src = hseq + ((size_t)(t - 1) * B + b0) * H;
  for (int i = tx; i < nb * h4; i += ntx) {
    const int r = i / h4, k4 = i % h4;
    sn_f4 v;
    SN_LD4_CG(src + (size_t)r * H + k4 * 4, v);
    sn_stv4(hs + (size_t)r * (H + 4) + k4 * 4, v);
  }
}

SN_HD void sn_lstm_seq_compute(const float* xg, const float* wsm, const float* hs, float* hseq, float* c, int t, int b0, int nb, int B,
                               int H, int bx, int tx) {
  const int u = tx >> 5, lane = tx & 31;
  const int j = bx * SN_PU + u;
  if (lane >= nb || j >= H) return;
  const int b = b0 + lane;
  const sn_f4 g = sn_ld4(xg + (((size_t)t * B + b) * H + j) * 4);
  float acc[4] = {g.v[0], g.v[1], g.v[2], g.v[3]};
  if (t > 0) {
    const float* hr = hs + (size_t)lane * (H + 4);
    const float* wu = wsm + (size_t)u * H * 4;
    for (int k = 0; k < H; k += 4) {
      const sn_f4 hv = sn_ldv4(hr + k);            // conflict-free: rows are H + 4 floats apart
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const sn_f4 w4 = sn_ldv4(wu + (size_t)(k + kk) * 4);   // same address for the whole warp: a broadcast
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(hv.v[kk], w4.v[q], acc[q]);
      }
    }
  }
  const size_t o = (size_t)b * H + j;
  const float cp = t > 0 ? c[o] : 0.f;
  const float cn = sn_sigmoid(acc[1]) * cp + sn_sigmoid(acc[0]) * tanhf(acc[2]);
  c[o] = cn;
  hseq[((size_t)t * B) * H + o] = sn_sigmoid(acc[3]) * tanhf(cn);
}
