// Residual vector quantisation (EnCodec RVQ, 8 x 1024 x 128 at 6 kbps): nearest-codebook encode and
// code -> latent embedding sum.
//
// Encode restates encodec's EuclideanCodebook.quantize: dist_j = -(|r|^2 - 2 r.e_j + |e_j|^2), argmax over j,
// residual r -= e[idx], 8 stages.  The decisive comparison is done in exact fp32 (no TF32 / bf16): the dot
// product is accumulated with fmaf in ascending d order and |r|^2, |e|^2 likewise, which is the order the
// C oracle (oracle/rvq_oracle.c) uses, so codes are bit-identical to the oracle; the first index wins ties.
// A tile of 128 frames stays in shared memory across all stages; codebooks (4 MB) stream from L2.
//
// Decode is a gather-sum: HBM-bound, 64 B of codes in and 512 B of fp32 latent out per frame.
#include "common.cuh"

namespace {

constexpr int TF = 128;   // frames per CTA
constexpr int TC = 128;   // codes per inner tile
constexpr int RD = 128;   // latent dimension (fixed by the tiling)

__global__ void __launch_bounds__(256, 1) rvq_encode_kernel(const float* __restrict__ lat, const float* __restrict__ cb, const float* __restrict__ cb_sq,
                                                            int64_t* __restrict__ codes, long long nframes, int T, int Q, int K) {
  extern __shared__ float sh[];
  float* sx = sh;                 // [RD][TF]   residual, d-major
  float* se = sh + RD * TF;       // [RD][TC]   codebook tile, d-major
  float* sxx = se + RD * TC;      // [TF]
  float* sbv = sxx + TF;          // [16][TF] best value per code-column group
  int* sbi = reinterpret_cast<int*>(sbv + 16 * TF);  // [16][TF]
  int* sidx = sbi + 16 * TF;      // [TF] winning index of the stage

  const int tid = threadIdx.x;
  const long long f0 = (long long)blockIdx.x * TF;

  // load the frame tile: lat[b, d, t] with frame f = b*T + t
  for (int i = tid; i < RD * TF; i += 256) {
    const int d = i / TF, fi = i % TF;
    const long long f = f0 + fi;
    float v = 0.f;
    if (f < nframes) {
      const long long b = f / T, t = f % T;
      v = lat[(b * RD + d) * T + t];
    }
    sx[i] = v;
  }
  __syncthreads();

  // frame group of this thread: frames tf*4 .. tf*4+3 and 64 + tf*4 .. 64 + tf*4+3.  Two 16-byte loads per d whose addresses are 16 bytes
  // apart across the 16 lanes of a frame group (contiguous 256 bytes = 2 wavefronts); the first version gave a thread 8 CONSECUTIVE
  // frames, i.e. 32-byte lane stride = 4-way bank conflicts on every residual load (ncu: 39 % of all shared wavefronts were conflicts).
  const int tf = tid % 16;
  const int tc = tid / 16;   // code group:  codes  tc*8 .. tc*8+7 of the tile
  auto fidx = [&](int i) { return (i < 4 ? 0 : 64 - 4) + tf * 4 + i; };   // frame of this thread's i-th accumulator row

  for (int q = 0; q < Q; ++q) {
    // |r|^2 per frame, ascending d with fmaf (one thread per frame)
    if (tid < TF) {
      float s = 0.f;
      for (int d = 0; d < RD; ++d) s = fmaf(sx[d * TF + tid], sx[d * TF + tid], s);
      sxx[tid] = s;
    }
    float best[8];
    int bidx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      best[i] = -INFINITY;
      bidx[i] = 0;
    }
    const float* cbq = cb + (long long)q * K * RD;
    for (int c0 = 0; c0 < K; c0 += TC) {
      __syncthreads();  // previous tile fully consumed (also orders sxx / sx updates)
      // codebook tile -> d-major shared
      for (int i = tid; i < TC * (RD / 4); i += 256) {
        const int c = i % TC, d4 = i / TC;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < K) v = *reinterpret_cast<const float4*>(cbq + (long long)(c0 + c) * RD + d4 * 4);
        se[(d4 * 4 + 0) * TC + c] = v.x;
        se[(d4 * 4 + 1) * TC + c] = v.y;
        se[(d4 * 4 + 2) * TC + c] = v.z;
        se[(d4 * 4 + 3) * TC + c] = v.w;
      }
      __syncthreads();
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
      for (int d = 0; d < RD; ++d) {
        const float4 xa = *reinterpret_cast<const float4*>(sx + d * TF + tf * 4);
        const float4 xb = *reinterpret_cast<const float4*>(sx + d * TF + 64 + tf * 4);
        const float4 ea = *reinterpret_cast<const float4*>(se + d * TC + tc * 8);
        const float4 eb = *reinterpret_cast<const float4*>(se + d * TC + tc * 8 + 4);
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv[i], ev[j], acc[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int code = c0 + tc * 8 + j;
        if (code < K) {
          const float ee = __ldg(cb_sq + (long long)q * K + code);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // dist = -((xx - 2*dot) + ee), the reference's evaluation order, each step rounded to fp32
            const float dist = -__fadd_rn(__fsub_rn(sxx[fidx(i)], __fmul_rn(2.f, acc[i][j])), ee);
            if (dist > best[i]) {  // ascending code order within the thread: strict > keeps the first maximum
              best[i] = dist;
              bidx[i] = code;
            }
          }
        }
      }
    }
    // reduce over the 16 code groups: larger value wins, equal values -> smaller index
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sbv[tc * TF + fidx(i)] = best[i];
      sbi[tc * TF + fidx(i)] = bidx[i];
    }
    __syncthreads();
    if (tid < TF) {
      float bv = sbv[tid];
      int bi = sbi[tid];
      for (int g = 1; g < 16; ++g) {
        const float v = sbv[g * TF + tid];
        const int ix = sbi[g * TF + tid];
        if (v > bv || (v == bv && ix < bi)) {
          bv = v;
          bi = ix;
        }
      }
      sidx[tid] = bi;
      const long long f = f0 + tid;
      if (f < nframes) {
        const long long b = f / T, t = f % T;
        codes[(b * Q + q) * T + t] = bi;
      }
    }
    __syncthreads();
    // residual update r -= e[idx]
    for (int i = tid; i < RD * TF; i += 256) {
      const int d = i / TF, fi = i % TF;
      sx[i] = __fsub_rn(sx[i], __ldg(cbq + (long long)sidx[fi] * RD + d));
    }
    __syncthreads();
  }
}

// |e|^2 per codebook entry, ascending d with fmaf
__global__ void rvq_cb_sq_kernel(const float* __restrict__ cb, float* __restrict__ out, long long n, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s = fmaf(cb[i * D + d], cb[i * D + d], s);
  out[i] = s;
}

// decode: one warp gathers the Q rows of a frame (float4 per lane), the CTA transposes 32 frames through
// shared memory so that the [B, D, T] output is written in 128-byte runs along T.
__global__ void rvq_decode_kernel(const int64_t* __restrict__ codes, const float* __restrict__ cb, float* __restrict__ lat, long long nframes,
                                  int T, int Q, int K) {
  __shared__ float tile[RD][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps
  const long long ntiles = (nframes + 31) / 32;
  for (long long tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
    const long long f0 = tb * 32;
    for (int fi = warp; fi < 32; fi += 8) {
      const long long f = f0 + fi;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < nframes) {
        const long long b = f / T, t = f % T;
        for (int q = 0; q < Q; ++q) {
          const long long c = min(max((long long)codes[(b * Q + q) * T + t], 0ll), (long long)K - 1);   // never read outside the codebook
          const float4 e = *reinterpret_cast<const float4*>(cb + ((long long)q * K + c) * RD + lane * 4);
          acc.x += e.x;
          acc.y += e.y;
          acc.z += e.z;
          acc.w += e.w;
        }
      }
      tile[lane * 4 + 0][fi] = acc.x;
      tile[lane * 4 + 1][fi] = acc.y;
      tile[lane * 4 + 2][fi] = acc.z;
      tile[lane * 4 + 3][fi] = acc.w;
    }
    __syncthreads();
    const long long f = f0 + lane;
    if (f < nframes) {
      const long long b = f / T, t = f % T;
      for (int d = warp; d < RD; d += 8) lat[(b * RD + d) * T + t] = tile[d][lane];
    }
    __syncthreads();
  }
}

// ---- decode, shared-memory variant.  The L2-gather kernel above moves 8 x 512 B of codebook rows per frame through L2 (4 KB per
// frame against 576 B of HBM traffic): it runs at the L2 bandwidth, not the HBM roofline.  Here a CTA owns a 4-float slice of the
// latent dimension and keeps that slice of ALL Q codebooks in shared memory (Q x K x 16 B = 128 KB for 8 x 1024); a thread owns a
// frame: 8 code loads (coalesced along t), 8 conflict-prone but on-chip LDS.128, the sum in q order (bit-equal to the sequential
// fp32 sum), 4 coalesced stores.  The int64 codes are narrowed to uint16 by a pre-pass, so the 32 slices re-read 16 B per frame
// from L2 instead of 64 B.
__global__ void rvq_pack_codes_kernel(const int64_t* __restrict__ codes, uint16_t* __restrict__ out, long long n, int K) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (uint16_t)min(max((long long)codes[i], 0ll), (long long)K - 1);
}

constexpr int DEC_THREADS = 1024;
constexpr int DEC_FPT = 4;      // frames per thread in flight: the loop is latency-bound (code loads from L2, then dependent LDS)
__global__ void __launch_bounds__(DEC_THREADS, 1) rvq_decode_smem_kernel(const uint16_t* __restrict__ codes, const float* __restrict__ cb,
                                                                        float* __restrict__ lat, long long nframes, int T, int Q, int K,
                                                                        int chunks) {
  extern __shared__ float4 scb[];        // [Q][K] : dims d0 .. d0+3 of every code
  const int slice = blockIdx.x % (RD / 4), chunk = blockIdx.x / (RD / 4);
  const int d0 = slice * 4;
  for (int i = threadIdx.x; i < Q * K; i += DEC_THREADS) scb[i] = __ldg(reinterpret_cast<const float4*>(cb + (long long)i * RD + d0));
  __syncthreads();
  const long long per = ((nframes + chunks - 1) / chunks + 31) / 32 * 32;      // whole warps of consecutive frames
  const long long f_begin = (long long)chunk * per, f_end = min(nframes, f_begin + per);
  for (long long f0 = f_begin + threadIdx.x; f0 < f_end; f0 += DEC_THREADS * DEC_FPT) {
    uint16_t c[DEC_FPT][8];
    long long ob[DEC_FPT];
    bool ok[DEC_FPT];
#pragma unroll
    for (int u = 0; u < DEC_FPT; ++u) {
      const long long f = f0 + (long long)u * DEC_THREADS;
      ok[u] = f < f_end;
      const long long fc = ok[u] ? f : f_begin;
      const long long b = fc / T;
      const int t = (int)(fc - b * T);
      ob[u] = (b * RD + d0) * T + t;
      const uint16_t* cp = codes + (b * Q) * T + t;
      if (Q == 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) c[u][q] = __ldg(cp + (long long)q * T);
      } else {
        for (int q = 0; q < 8; ++q) c[u][q] = q < Q ? __ldg(cp + (long long)q * T) : (uint16_t)0;
      }
    }
#pragma unroll
    for (int u = 0; u < DEC_FPT; ++u) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < Q) {      // q ascending, fp32 adds: bit-equal to the sequential codeword sum
          const float4 e = scb[q * K + c[u][q]];
          acc.x += e.x, acc.y += e.y, acc.z += e.z, acc.w += e.w;
        }
      }
      if (ok[u]) {
        float* o = lat + ob[u];
        __stcs(o, acc.x);
        __stcs(o + (long long)T, acc.y);
        __stcs(o + 2ll * T, acc.z);
        __stcs(o + 3ll * T, acc.w);
      }
    }
  }
}

__global__ void codes_affine_kernel(const int64_t* __restrict__ codes, float* __restrict__ x0, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // dataloader.py:64,77,168-170: Normalize(0.5, 0.5)(codes / 1023) = (c/1023 - 0.5) / 0.5, each step rounded to fp32
    const float u = __fdiv_rn((float)codes[i], 1023.f);
    x0[i] = __fdiv_rn(__fsub_rn(u, 0.5f), 0.5f);
  }
}
__global__ void codes_affine_inv_kernel(const float* __restrict__ x, int64_t* __restrict__ codes, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = rintf((x[i] + 1.f) * 511.5f);
    codes[i] = (int64_t)fminf(fmaxf(c, 0.f), 1023.f);
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int pt_rvq_cb_sq(const float* codebooks, float* out, int Q, int K, int D, void* stream) {
  PT_REQUIRE(Q > 0 && K > 0 && D > 0, "rvq_cb_sq: Q=%d K=%d D=%d", Q, K, D);
  const long long n = (long long)Q * K;
  rvq_cb_sq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST>>>(codebooks, out, n, D);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

// cb_sq: caller-provided [Q, K] fp32 scratch (filled here)
extern "C" int pt_rvq_encode_ws(const float* latents, const float* codebooks, float* cb_sq, int64_t* codes, int B, int D, int T, int Q, int K,
                                void* stream) {
  PT_REQUIRE(B > 0 && T > 0 && Q > 0 && K > 0, "rvq_encode: B=%d T=%d Q=%d K=%d", B, T, Q, K);
  PT_REQUIRE(D == RD, "rvq_encode: latent dimension must be %d (EnCodec), got %d", RD, D);
  if (int r = pt_rvq_cb_sq(codebooks, cb_sq, Q, K, D, stream)) return r;
  const size_t smem = sizeof(float) * (RD * TF + RD * TC + TF + 16 * TF) + sizeof(int) * (16 * TF + TF);
  PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nframes = (long long)B * T;
  rvq_encode_kernel<<<(unsigned)((nframes + TF - 1) / TF), 256, smem, ST>>>(latents, codebooks, cb_sq, codes, nframes, T, Q, K);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_rvq_decode(const int64_t* codes, const float* codebooks, float* latents, int B, int D, int T, int Q, int K, void* stream) {
  PT_REQUIRE(B > 0 && T > 0 && Q > 0 && K > 0, "rvq_decode: B=%d T=%d Q=%d K=%d", B, T, Q, K);
  PT_REQUIRE(D == RD, "rvq_decode: latent dimension must be %d (EnCodec), got %d", RD, D);
  const long long nframes = (long long)B * T;
  long long blocks = (nframes + 31) / 32;
  const long long cap = 16ll * pt_num_sms();
  if (blocks > cap) blocks = cap;
  rvq_decode_kernel<<<(unsigned)blocks, 256, 0, ST>>>(codes, codebooks, latents, nframes, T, Q, K);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

// scratch: caller-provided, B * Q * T uint16 (the narrowed codes)
extern "C" int pt_rvq_decode_ws(const int64_t* codes, const float* codebooks, float* latents, void* scratch, int B, int D, int T, int Q, int K,
                                void* stream) {
  PT_REQUIRE(B > 0 && T > 0 && Q > 0 && K > 0, "rvq_decode: B=%d T=%d Q=%d K=%d", B, T, Q, K);
  PT_REQUIRE(D == RD, "rvq_decode: latent dimension must be %d (EnCodec), got %d", RD, D);
  const size_t smem = (size_t)Q * K * sizeof(float4);
  if (scratch == nullptr || smem > 200 * 1024 || K > 65536 || Q > 8) return pt_rvq_decode(codes, codebooks, latents, B, D, T, Q, K, stream);
  const long long nframes = (long long)B * T, ncodes = nframes * Q;
  long long pb = (ncodes + 255) / 256;
  if (pb > 8ll * pt_num_sms()) pb = 8ll * pt_num_sms();
  rvq_pack_codes_kernel<<<(unsigned)pb, 256, 0, ST>>>(codes, (uint16_t*)scratch, ncodes, K);
  PT_LAUNCH_CHECK();
  PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(rvq_decode_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // one CTA per SM: 32 slices x `chunks` frame ranges; every CTA pays one 128 KB codebook-slice load, so give it >= 4 K frames
  int chunks = pt_num_sms() / (RD / 4);
  if (chunks < 1) chunks = 1;
  while (chunks > 1 && nframes / chunks < 4096) --chunks;
  rvq_decode_smem_kernel<<<(unsigned)(chunks * (RD / 4)), DEC_THREADS, smem, ST>>>((const uint16_t*)scratch, codebooks, latents, nframes, T, Q, K,
                                                                                   chunks);
  PT_LAUNCH_CHECK();
  return PT_OK;
}

extern "C" int pt_codes_affine(const int64_t* codes, float* x0, int64_t n, void* stream) {
  PT_REQUIRE(n > 0, "codes_affine: n=%lld", (long long)n);
  long long b = (n + 255) / 256;
  const long long cap = 8ll * pt_num_sms();
  if (b > cap) b = cap;
  codes_affine_kernel<<<(unsigned)b, 256, 0, ST>>>(codes, x0, n);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
extern "C" int pt_codes_affine_inv(const float* x, int64_t* codes, int64_t n, void* stream) {
  PT_REQUIRE(n > 0, "codes_affine_inv: n=%lld", (long long)n);
  long long b = (n + 255) / 256;
  const long long cap = 8ll * pt_num_sms();
  if (b > cap) b = cap;
  codes_affine_inv_kernel<<<(unsigned)b, 256, 0, ST>>>(x, codes, n);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
