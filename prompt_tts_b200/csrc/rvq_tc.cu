// RVQ nearest-codebook quantise with tensor-core PRE-SELECTION and exact fp32 RE-RANKING.
//
// The reference decision (encodec EuclideanCodebook.quantize, restated in oracle/rvq_oracle.c) is the argmax over 1024 codes of
//     dist_j = -((|r|^2 - 2 r.e_j) + |e_j|^2)        every step rounded to fp32, r.e_j an fmaf chain in ascending d,
// first index on ties.  Evaluating all 1024 chains per frame and stage is 2.1 MFLOP per frame on the FMA pipe (csrc/rvq.cu: 35 of
// 72 TFLOP/s).  Here the tensor cores compute an APPROXIMATE score for all codes and the exact chain runs only for the codes that can
// still be the maximum:
//     a_j = 2 * dot_bf16(r, e_j) - |e_j|^2          dot_bf16: tcgen05.mma on bf16 roundings of r and e, fp32 accumulation in TMEM
//     |a_j - (dist_j + |r|^2)| <= Bnd = 0.0162 |r| max_j|e_j| + 5e-7 (|r|^2 + max|e|^2 + 2 |r| max|e|)
// (bf16 keeps 8 significant bits: unit roundoff 2^-8 per operand, 2^-7 + 2^-16 = 0.00783 relative on every product, Cauchy-Schwarz on
// the sum, x2 for the factor 2 = 0.01566, + 3 % for the tensor core's accumulation; the second term covers the fp32 roundings of both
// evaluations), hence   argmax_j dist_j  is in  { j : a_j >= max_j a_j - 2 Bnd }.
// For Gaussian data that set has ~2 members; each member is re-evaluated with EXACTLY the oracle's arithmetic in ascending j with
// a strict comparison, so the codes are bit-identical to the oracle's (and to the fp32 kernel's) by construction -- the tensor cores
// only decide what is NOT evaluated.  A frame whose candidate list overflows (degenerate codebooks) is scanned exhaustively.
//
// One persistent CTA per SM, 128 frames per work item, 320 threads:
//   warp 0     TMA producer: bf16 codebook tiles (128 codes x 128 d, SWIZZLE_128B, two 64-d blocks) into a 3-stage ring
//   warp 1     MMA issuer:   per tile 8 x tcgen05.mma.cta_group::1.kind::f16 (M = 128 frames, N = 128 codes, K = 16) into one of
//              four 128-column TMEM accumulators
//   warps 2-9  select: two threads per frame.  Keep the fp32 residual row in shared memory, writes its bf16 copy as the A operand (in
//              the swizzle the MMA expects), streams the scores out of TMEM (tcgen05.ld), keeps the near-maximal ones, re-ranks
//              them exactly, writes the code and subtracts the chosen codeword (fp32, exact) -- 8 stages, residual never leaves
//              the SM.
#include <math.h>

#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int RD = 128;          // latent dimension
constexpr int TFR = 128;         // frames per work item (MMA M)
constexpr int TN = 128;          // codes per tile (MMA N)
constexpr int NSTAGE = 3;        // codebook tile ring
constexpr int NACC = 4;          // TMEM accumulators (4 x 128 columns)
constexpr int CAP = 32;          // candidate slots per frame
constexpr int RSTRIDE = RD + 1;  // fp32 residual row stride (conflict-free for thread-per-row AND lane-per-column access)

constexpr int A_BYTES = TFR * RD * 2;            // 32 KB: two 64-d blocks of 16 KB
constexpr int B_BYTES = TN * RD * 2;             // 32 KB per stage
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + A_BYTES;
constexpr int OFF_R = OFF_B + NSTAGE * B_BYTES;
constexpr int OFF_EE = OFF_R + TFR * RSTRIDE * 4;
constexpr int MAXK = 1024;
constexpr int OFF_CV = OFF_EE + MAXK * 4;        // candidate values [CAP][TFR] fp32
constexpr int OFF_CI = OFF_CV + CAP * TFR * 4;   // candidate indices [CAP][TFR] u16
constexpr int OFF_XX = OFF_CI + CAP * TFR * 2;   // |r|^2 per frame (read by whichever lane re-ranks one of the frame's codes)
constexpr int OFF_PRE = OFF_XX + TFR * 4;        // per select warp: 33 (+3) prefix sums of candidate counts
constexpr int OFF_MX = OFF_PRE + 8 * 36 * 4;     // per half: running maximum [2][TFR] fp32
constexpr int OFF_CN = OFF_MX + 2 * TFR * 4;     // per half: surviving candidates [2][TFR] int
constexpr int OFF_OV = OFF_CN + 2 * TFR * 4;     // per half: list overflow flag [2][TFR] int
constexpr int OFF_BI = OFF_OV + 2 * TFR * 4;     // chosen code per frame [TFR] int
constexpr int OFF_BAR = OFF_BI + TFR * 4;
constexpr int CAPH = CAP / 2;                    // candidate slots per frame and half
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");

struct alignas(64) TcParams {
  CUtensorMap tmB;           // bf16 codebooks as (128 d, Q*K codes)
  const float* lat;          // [B, 128, T]
  const float* cb;           // [Q, K, 128] fp32
  const float* cb_sq;        // [Q, K]   |e|^2 (fmaf chain, ascending d)
  const float* emax;         // [Q]      max_j |e_j|
  int64_t* codes;            // [B, Q, T]
  long long nframes;
  int T, Q, K;
  unsigned long long* overflow_count;   // statistics (frames that took the exhaustive path); may be null
  float* dbg_scores;         // diagnostic: approximate scores a_j of stage 0 of the first 128 frames, [128][K]; normally null
};

// exact reference distance of code j for the residual row `rr` (|r|^2 = xx): the oracle's arithmetic, step by step
__device__ __forceinline__ float exact_dist(const float* __restrict__ rr, float xx, const float* __restrict__ e, float ee) {
  float dot = 0.f;
#pragma unroll 8
  for (int d4 = 0; d4 < RD / 4; ++d4) {
    const float4 ev = __ldg(reinterpret_cast<const float4*>(e) + d4);
    dot = fmaf(rr[d4 * 4 + 0], ev.x, dot);
    dot = fmaf(rr[d4 * 4 + 1], ev.y, dot);
    dot = fmaf(rr[d4 * 4 + 2], ev.z, dot);
    dot = fmaf(rr[d4 * 4 + 3], ev.w, dot);
  }
  return -__fadd_rn(__fsub_rn(xx, __fmul_rn(2.f, dot)), ee);
}

__global__ void __launch_bounds__(320, 1) rvq_encode_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sA = sbase + OFF_A, sB = sbase + OFF_B, bar0 = sbase + OFF_BAR;
  float* sr = reinterpret_cast<float*>(sgen + OFF_R);
  float* see = reinterpret_cast<float*>(sgen + OFF_EE);
  float* scv = reinterpret_cast<float*>(sgen + OFF_CV);
  uint16_t* sci = reinterpret_cast<uint16_t*>(sgen + OFF_CI);
  float* sxx = reinterpret_cast<float*>(sgen + OFF_XX);
  float* smx = reinterpret_cast<float*>(sgen + OFF_MX);
  int* scn = reinterpret_cast<int*>(sgen + OFF_CN);
  int* sov = reinterpret_cast<int*>(sgen + OFF_OV);
  int* sbi = reinterpret_cast<int*>(sgen + OFF_BI);
  auto b_full = [&](int s) { return bar0 + 8u * s; };
  auto b_empty = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto acc_full = [&](int a) { return bar0 + 8u * (2 * NSTAGE + a); };
  auto acc_empty = [&](int a) { return bar0 + 8u * (2 * NSTAGE + NACC + a); };
  const uint32_t a_ready = bar0 + 8u * (2 * NSTAGE + 2 * NACC);
  const uint32_t tmem_slot = bar0 + 8u * (2 * NSTAGE + 2 * NACC + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_k = p.K / TN;                      // code tiles per stage
  const long long n_items = (p.nframes + TFR - 1) / TFR;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 8);      // one elected lane per select warp
    }
    mbar_init(a_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + OFF_BAR + 8 * (2 * NSTAGE + 2 * NACC + 1));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 1;
    for (long long w = blockIdx.x; w < n_items; w += gridDim.x)
      for (int q = 0; q < p.Q; ++q)
        for (int n = 0; n < ntile_k; ++n) {
          mbar_wait(b_empty(s), ph);
          if (leader) {
            mbar_expect_tx(b_full(s), B_BYTES);
            const uint32_t dst = sB + s * B_BYTES;
            const int row = q * p.K + n * TN;
            tma_load_4d(dst, &p.tmB, b_full(s), 0, row, 0, 0);
            tma_load_4d(dst + TN * 128, &p.tmB, b_full(s), 64, row, 0, 0);
          }
          if (++s == NSTAGE) s = 0, ph ^= 1;
        }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t idesc = idesc_f16(0, 0, TN, TFR);
    const uint64_t dA = umma_desc(sA, 0, 1024);
    const uint64_t dB = umma_desc(sB, 0, 1024);
    int s = 0;
    uint32_t ph = 0;
    uint32_t g = 0;        // tiles issued so far (accumulator = g & 3)
    uint32_t sg = 0;       // stages started so far
    for (long long w = blockIdx.x; w < n_items; w += gridDim.x)
      for (int q = 0; q < p.Q; ++q, ++sg) {
        mbar_wait(a_ready, sg & 1);      // this stage's bf16 residual tile is in shared memory (and visible to the async proxy)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int n = 0; n < ntile_k; ++n, ++g) {
          const uint32_t a = g & (NACC - 1);
          mbar_wait(acc_empty(a), ((g / NACC) & 1) ^ 1);
          mbar_wait(b_full(s), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (leader) {
            const uint64_t bd = dB + (uint64_t)(s * (B_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < RD / 16; ++k) {
              const uint64_t off_a = (uint64_t)((k >> 2) * ((TFR * 128) >> 4) + (k & 3) * 2);
              const uint64_t off_b = (uint64_t)((k >> 2) * ((TN * 128) >> 4) + (k & 3) * 2);
              umma_f16(tmem + a * TN, dA + off_a, bd + off_b, idesc, k > 0 ? 1u : 0u);
            }
            umma_commit(b_empty(s));
            umma_commit(acc_full(a));
          }
          __syncwarp();
          if (++s == NSTAGE) s = 0, ph ^= 1;
        }
      }
  } else {
    // ------------------------------------------------------------------ select warps.  Two threads per frame ("halves": warps 2-5
    // and 6-9, the same TMEM lane quarter each): half h examines the 32-column chunks with (chunk & 1) == h, keeps its own running
    // maximum and candidate list, and re-ranks its own candidates; the halves meet once per stage.  Everything in this loop is bound
    // by dependent latencies (TMEM load -> FMA -> max tree -> compare; global codeword loads), so the second warp per scheduler and
    // the halved per-thread work are what sets the speed.
    const int qd = warp & 3;                       // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;
    const int f = qd * 32 + lane;                  // frame (row) of this thread inside the work item
    const int st = threadIdx.x - 64;               // 0..255 among the select threads
    const uint32_t tl = tmem + ((uint32_t)(qd * 32) << 16);
    float* rr = sr + f * RSTRIDE;
    int* spre = reinterpret_cast<int*>(sgen + OFF_PRE) + (warp - 2) * 36;      // per-warp prefix sums of the candidate counts
    float* lv = scv + half * (CAPH * TFR);         // this half's candidate values / indices: [CAPH][TFR]
    uint16_t* li = sci + half * (CAPH * TFR);
    uint32_t g = 0, sg = 0;
    for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
      const long long fr = w * TFR + f;
      const bool live = fr < p.nframes;
      long long b = 0;
      int t = 0;
      if (live) {
        b = fr / p.T;
        t = (int)(fr - b * p.T);
      }
      // residual tile: lat[b, d, t], thread = (frame, half of d) (coalesced along t); 16 loads in flight per thread
      {
        const float* lp = p.lat + (b * RD) * p.T + t;
#pragma unroll 1
        for (int d0 = half * (RD / 2); d0 < (half + 1) * (RD / 2); d0 += 16) {
          float tmp[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) tmp[u] = live ? __ldg(lp + (long long)(d0 + u) * p.T) : 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) rr[d0 + u] = tmp[u];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");        // both halves of every row are in shared memory
      for (int q = 0; q < p.Q; ++q, ++sg) {
        // |r|^2: ascending d, fmaf (the oracle's order) -- both halves compute it (each needs the window)
        float xx = 0.f;
#pragma unroll 8
        for (int d = 0; d < RD; ++d) xx = fmaf(rr[d], rr[d], xx);
        // bf16 copy of the row -> A operand tile (K-major, SWIZZLE_128B: 16-byte chunk c of row f lives at chunk c ^ (f & 7));
        // half h writes the 64-d block h
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int c = half * 8 + cc;
          uint4 v;
          v.x = f2_to_bf2(rr[c * 8 + 0], rr[c * 8 + 1]);
          v.y = f2_to_bf2(rr[c * 8 + 2], rr[c * 8 + 3]);
          v.z = f2_to_bf2(rr[c * 8 + 4], rr[c * 8 + 5]);
          v.w = f2_to_bf2(rr[c * 8 + 6], rr[c * 8 + 7]);
          *reinterpret_cast<uint4*>(sgen + OFF_A + half * (TFR * 128) + f * 128 + ((cc ^ (f & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        // |e|^2 of this stage -> shared (everyone has finished reading the previous stage's table: they passed the barrier below)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {
          float tmp[MAXK / 256];
#pragma unroll
          for (int u = 0; u < MAXK / 256; ++u) tmp[u] = (st + 256 * u < p.K) ? __ldg(p.cb_sq + (long long)q * p.K + st + 256 * u) : 0.f;
#pragma unroll
          for (int u = 0; u < MAXK / 256; ++u)
            if (st + 256 * u < p.K) see[st + 256 * u] = tmp[u];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);

        const float rn = sqrtf(xx), em = __ldg(p.emax + q);
        const float win = 2.f * (0.0162f * rn * em + 5e-7f * (xx + em * em + 2.f * rn * em)) + 1e-30f;   // 2 * Bnd
        // ---- stream the approximate scores of this half's chunks.  The maxima of groups of 4 fall out of the max tree; only a
        // group whose maximum is inside the window of the running maximum is looked at code by code (the final threshold can only
        // be higher than the running one).
        float m = -INFINITY;
        int cnt = 0;
        bool overflow = false;
        const int ntot = ntile_k * (TN / 32);      // 32-column chunks of this stage; this half takes chunk half, half + 2, ...
        const bool dbg = p.dbg_scores != nullptr && w == 0 && q == 0;
        uint32_t va[32], vb[32];                   // two chunks in flight; explicit arrays keep them in registers
        mbar_wait(acc_full(g & (NACC - 1)), (g / NACC) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32(tl + (g & (NACC - 1)) * TN + half * 32, va);
        auto chunk = [&](int ch, const uint32_t* vv, uint32_t* vnext) {
          const int n = ch >> 2, c = ch & 3;
          const uint32_t a = (g + n) & (NACC - 1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c >= 2) {      // this half's last chunk of the tile is in registers: hand the accumulator back (one arrival per warp)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(a));
          }
          if (ch + 2 < ntot) {      // this half's next chunk flies while the current one is examined
            const int n2 = (ch + 2) >> 2, c2 = (ch + 2) & 3;
            const uint32_t a2 = (g + n2) & (NACC - 1);
            if (c2 < 2) {
              mbar_wait(acc_full(a2), ((g + n2) / NACC) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            tmem_ld32(tl + a2 * TN + c2 * 32, vnext);
          }
          const int j0 = n * TN + c * 32;
          float sc[32], gm[8];
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 e4 = *reinterpret_cast<const float4*>(see + j0 + i);
            sc[i + 0] = fmaf(2.f, __uint_as_float(vv[i + 0]), -e4.x);
            sc[i + 1] = fmaf(2.f, __uint_as_float(vv[i + 1]), -e4.y);
            sc[i + 2] = fmaf(2.f, __uint_as_float(vv[i + 2]), -e4.z);
            sc[i + 3] = fmaf(2.f, __uint_as_float(vv[i + 3]), -e4.w);
            gm[i >> 2] = fmaxf(fmaxf(sc[i], sc[i + 1]), fmaxf(sc[i + 2], sc[i + 3]));
            cm = fmaxf(cm, gm[i >> 2]);
          }
          if (dbg) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.dbg_scores[(long long)f * p.K + j0 + i] = sc[i];
          }
          m = fmaxf(m, cm);
          const float thr = m - win;
          if (cm >= thr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (gm[k] >= thr) {      // (a non-inlined push per group was measured: 54.7 vs 61.4 M frames/s -- the predicated form wins)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const bool hit = sc[4 * k + u] >= thr;
                  const bool room = cnt < CAPH;
                  if (hit && room) {
                    lv[cnt * TFR + f] = sc[4 * k + u];
                    li[cnt * TFR + f] = (uint16_t)(j0 + 4 * k + u);
                  }
                  overflow |= hit && !room;
                  cnt += (hit && room) ? 1 : 0;
                }
              }
            }
          }
          if (cnt > CAPH / 2) {                    // drop what the running threshold has already ruled out (keeps ascending order)
            int k2 = 0;
            for (int k = 0; k < cnt; ++k) {
              const float cv = lv[k * TFR + f];
              const uint16_t ci = li[k * TFR + f];
              if (cv >= thr) {
                lv[k2 * TFR + f] = cv;
                li[k2 * TFR + f] = ci;
                ++k2;
              }
            }
            cnt = k2;
          }
        };
        for (int ch = half; ch < ntot; ch += 4) {  // K % 128 == 0: every half has an even number of chunks
          chunk(ch, va, vb);
          chunk(ch + 2, vb, va);
        }
        g += ntile_k;
        // ---- the halves meet: the frame's maximum is the larger of the two running maxima
        smx[half * TFR + f] = m;
        sov[half * TFR + f] = overflow ? 1 : 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mfin = fmaxf(smx[f], smx[TFR + f]);
        const bool ovf_any = (sov[f] | sov[TFR + f]) != 0;
        const float* cbq = p.cb + (long long)q * p.K * RD;
        {      // final filter of this half's list (in place, ascending order kept)
          const float thr = mfin - win;
          int k2 = 0;
          for (int k = 0; k < cnt; ++k) {
            const float cv = lv[k * TFR + f];
            const uint16_t ci = li[k * TFR + f];
            if (cv >= thr) {
              lv[k2 * TFR + f] = cv;
              li[k2 * TFR + f] = ci;
              ++k2;
            }
          }
          cnt = ovf_any ? 0 : k2;
        }
        // ---- exact re-ranking, balanced over the warp: this warp's surviving codes (~1 per frame and half) are dealt round-robin to
        // its lanes whatever frame they belong to; a lane evaluates a code with the oracle's arithmetic (fmaf chain over ascending
        // d) and leaves the exact distance in the code's slot.
        {
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
          }
          spre[lane + 1] = incl;
          if (lane == 0) spre[0] = 0;
          if (half == 0) sxx[f] = xx;
          scn[half * TFR + f] = cnt;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int total = spre[32];
          for (int it = lane; it < total; it += 32) {
            int lo = 0, hi = 31;                   // owner lane: the last one whose exclusive prefix is <= it
#pragma unroll
            for (int bs = 0; bs < 5; ++bs) {
              const int mid = (lo + hi + 1) >> 1;
              if (spre[mid] <= it) lo = mid; else hi = mid - 1;
            }
            const int fo = qd * 32 + lo, slot = it - spre[lo];
            const int j = li[slot * TFR + fo];
            lv[slot * TFR + fo] = exact_dist(sr + fo * RSTRIDE, sxx[fo], cbq + (long long)j * RD, see[j]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // ---- the owner (half 0) takes the maximum over both halves' slots; on equal distances the smaller index wins (= the first
        // maximum of the oracle's ascending scan)
        int bi = 0;
        if (half == 0) {
          float best = -INFINITY;
          bi = 0x7fffffff;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c2 = scn[hh * TFR + f];
            for (int k = 0; k < c2; ++k) {
              const float dq = scv[(hh * CAPH + k) * TFR + f];
              const int jq = sci[(hh * CAPH + k) * TFR + f];
              if (dq > best || (dq == best && jq < bi)) {
                best = dq;
                bi = jq;
              }
            }
          }
          // degenerate cases (more near-maximal codes than slots): exhaustive exact scan of the frame, by the whole warp -- lane l takes
          // codes l, l + 32, ... in ascending order (strict >), then the lanes' bests are merged with the smaller index winning ties
          unsigned om = __ballot_sync(0xffffffffu, ovf_any);
          while (om) {
            const int fo = __ffs(om) - 1;
            om &= om - 1;
            if (lane == 0 && p.overflow_count) atomicAdd(p.overflow_count, 1ull);
            const float* ro = sr + (qd * 32 + fo) * RSTRIDE;
            const float xo = sxx[qd * 32 + fo];
            float lb = -INFINITY;
            int lj = 0x7fffffff;
            for (int j = lane; j < p.K; j += 32) {
              const float dist = exact_dist(ro, xo, cbq + (long long)j * RD, see[j]);
              if (dist > lb) lb = dist, lj = j;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, lb, o);
              const int oj = __shfl_xor_sync(0xffffffffu, lj, o);
              if (ob > lb || (ob == lb && oj < lj)) lb = ob, lj = oj;
            }
            if (lane == fo) best = lb, bi = lj;
          }
          if (live) p.codes[(b * p.Q + q) * p.T + t] = bi;
          sbi[f] = bi;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // ---- residual update r -= e[bi] (fp32, exact): each warp takes 16 of its quarter's frames (half h: frames 16 h .. 16 h + 15)
        // with coalesced codeword loads, four frames (16 loads per lane) in flight
#pragma unroll 1
        for (int fi = half * 16; fi < half * 16 + 16; fi += 4) {
          float ev[4][RD / 32];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float* e = cbq + (long long)sbi[qd * 32 + fi + u] * RD;
#pragma unroll
            for (int i = 0; i < RD / 32; ++i) ev[u][i] = __ldg(e + lane + 32 * i);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float* ro = sr + (qd * 32 + fi + u) * RSTRIDE;
#pragma unroll
            for (int i = 0; i < RD / 32; ++i) ro[lane + 32 * i] = __fsub_rn(ro[lane + 32 * i], ev[u][i]);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");        // updated rows are visible to both halves of every frame
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// bf16 copy of the codebooks, |e|^2 (fmaf chain, ascending d -- the oracle's) and max_j |e_j| per stage.  One warp per code.
__global__ void rvq_tc_prep_kernel(const float* __restrict__ cb, bf16* __restrict__ cb16, float* __restrict__ cb_sq, unsigned int* __restrict__ emax_bits,
                                   int Q, int K) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Q * K) return;
  const float* e = cb + i * RD;
  float s = 0.f;
  for (int d = 0; d < RD; d += 8) {
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      f[u] = e[d + u];
      s = fmaf(f[u], f[u], s);
    }
    store8(cb16 + i * RD + d, f);
  }
  cb_sq[i] = s;
  atomicMax(emax_bits + (int)(i / K), __float_as_uint(sqrtf(s) * 1.0000002f));      // non-negative floats order like their bit patterns
}

}  // namespace

#define ST ((cudaStream_t)stream)

static float* g_dbg_scores = nullptr;
// diagnostic hook (tests / tools only): subsequent pt_rvq_encode_tc launches dump the approximate stage-0 scores of their first 128
// frames into `buf` ([128][K] fp32); pass NULL to switch it off
extern "C" int pt_rvq_tc_debug_scores(float* buf) {
  g_dbg_scores = buf;
  return PT_OK;
}

extern "C" size_t pt_rvq_encode_tc_scratch_bytes(int Q, int K) {
  return (size_t)Q * K * 4 + 256 + (size_t)Q * K * RD * 2 + 64;
}

// scratch layout: [Q*K fp32 |e|^2][64 x u32 emax][Q*K*128 bf16 codebooks][u64 overflow counter]
extern "C" int pt_rvq_encode_tc(const float* latents, const float* codebooks, void* scratch, int prep, int64_t* codes, int B, int D, int T, int Q,
                                int K, void* stream) {
  PT_REQUIRE(B > 0 && T > 0 && Q > 0 && Q <= 64 && K > 0, "rvq_encode_tc: B=%d T=%d Q=%d K=%d", B, T, Q, K);
  PT_REQUIRE(D == RD, "rvq_encode_tc: latent dimension must be %d (EnCodec), got %d", RD, D);
  PT_REQUIRE(K % TN == 0 && K <= MAXK, "rvq_encode_tc: K=%d must be a multiple of %d and <= %d (use pt_rvq_encode_ws otherwise)", K, TN, MAXK);
  PT_REQUIRE(scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "rvq_encode_tc: scratch must be 256-byte aligned");
  uint8_t* sp = reinterpret_cast<uint8_t*>(scratch);
  float* cb_sq = reinterpret_cast<float*>(sp);
  unsigned int* emax = reinterpret_cast<unsigned int*>(sp + (size_t)Q * K * 4);
  bf16* cb16 = reinterpret_cast<bf16*>(sp + (size_t)Q * K * 4 + 256);
  unsigned long long* ovf = reinterpret_cast<unsigned long long*>(sp + (size_t)Q * K * 4 + 256 + (size_t)Q * K * RD * 2);
  if (prep) {
    PT_CUDA_OK(cudaMemsetAsync(emax, 0, 256, ST));
    PT_CUDA_OK(cudaMemsetAsync(ovf, 0, 8, ST));
    const long long n = (long long)Q * K;
    rvq_tc_prep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ST>>>(codebooks, cb16, cb_sq, emax, Q, K);
    PT_LAUNCH_CHECK();
  }
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  EncodeTiledFn fn = get_encode_fn();
  PT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  {
    cuuint64_t dims[4] = {(cuuint64_t)RD, (cuuint64_t)Q * K, 1, 1};
    cuuint64_t strides[3] = {(cuuint64_t)RD * 2, (cuuint64_t)RD * 2 * (cuuint64_t)Q * K, (cuuint64_t)RD * 2 * (cuuint64_t)Q * K};
    cuuint32_t box[4] = {64, (cuuint32_t)TN, 1, 1}, estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&tp.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, cb16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pt_set_error("rvq_encode_tc: cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
      return PT_ECUDA;
    }
  }
  tp.lat = latents;
  tp.cb = codebooks;
  tp.cb_sq = cb_sq;
  tp.emax = reinterpret_cast<const float*>(emax);
  tp.codes = codes;
  tp.nframes = (long long)B * T;
  tp.T = T;
  tp.Q = Q;
  tp.K = K;
  tp.overflow_count = ovf;
  tp.dbg_scores = g_dbg_scores;
  PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(rvq_encode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const long long items = (tp.nframes + TFR - 1) / TFR;
  const int sms = pt_num_sms();
  rvq_encode_tc_kernel<<<(unsigned)(items < sms ? items : sms), 320, SMEM_BYTES, ST>>>(tp);
  PT_LAUNCH_CHECK();
  return PT_OK;
}
