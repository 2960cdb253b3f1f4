// Fused softmax attention (forward and backward) on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces diffusers' AttnProcessor2_0 (F.scaled_dot_product_attention, no mask, no dropout) as the reference
// uses it from tts/ldm/transformer_1d.py:258-265 and tts/models.py:95-100,117-119.  No [Lq, Lk] matrix ever
// reaches HBM: logits live in TMEM; probabilities (and dS) are written back into TMEM over the logits they came from and feed
// the second MMA as its A operand from tensor memory -- they never touch shared memory either.
//
// One kernel template, three modes.  A CTA owns 128 "resident" rows and streams 64-row tiles of the other side:
//
//   mode  resident (R1,R2)  streamed (S1,S2)  stage 1 (TMEM X)              transform (CUDA cores)       stage 2 (TMEM ACC)
//   FWD   Q                 K, V              X1 = Q K^T                    P  = exp(X1*s - m)           O  += P V          (/ rowsum)
//   DQ    Q, dO             K, V              X1 = Q K^T,  X2 = dO V^T      dS = P (X2 - D) s            dQ += dS K
//   DKV   K, V              Q, dO             X1 = K Q^T,  X2 = V dO^T      P^T, dS^T                    dV += P^T dO ; dK += dS^T Q
//
// with P = exp(X1*s - lse) in the backward modes (lse saved by FWD, D = rowsum(dO * O) from a small pre-pass).
// FWD is a single-sweep online softmax: each transform group keeps its own running row maximum, row sum and accumulator; the
// reference maximum only moves when a row's maximum grows by more than 2^8 (then that group's accumulator is rescaled in TMEM,
// which is rare), and the two groups are merged exactly in the epilogue.
//
// Warp roles (352 threads): warp 0 TMA producer, warp 1 stage-1 MMA issuer, warp 10 stage-2 MMA issuer, warps 2-5 and 6-9 two
// transform groups that ping-pong over the streamed tiles (group g owns TMEM buffer X[g]),
// so the tensor core computes the logits of tile j+1 while the CUDA cores exponentiate tile j.
// Every streamed tile is used twice from the same shared-memory bytes: K-major as the B operand of stage 1 and
// MN-major as the B operand of stage 2 (only the UMMA descriptor differs).
#include <math.h>

#include <type_traits>

#ifdef PT_ATTN_TRACE
// development builds: a wait that has spun 2^20 times writes (barrier offset, parity) of its warp and the raw words of all barriers of
// its CTA to g_watch (pinned host memory: survives the trap that follows).  Layout per CTA: [12 warps] + [40 barrier words].
#include <stdint.h>
__device__ unsigned long long* g_watch = nullptr;
__device__ __forceinline__ void pt_mbar_watch(uint32_t bar, uint32_t parity) {
  if (g_watch == nullptr) return;
  volatile unsigned long long* w = g_watch + (size_t)blockIdx.x * 52;
  const uint32_t base = bar & ~511u;
  w[threadIdx.x >> 5] = (1ull << 63) | ((unsigned long long)(threadIdx.x & 31) << 40) | ((unsigned long long)parity << 32) | (bar - base);
  for (int i = 0; i < 40; ++i) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(base + 8u * i));
    w[12 + i] = v;
  }
}
#define PT_MBAR_WATCH(bar, parity) pt_mbar_watch(bar, parity)
#endif
#include "tc_common.cuh"

namespace {
using namespace tc;

enum { MODE_FWD = 0, MODE_DQ = 1, MODE_DKV = 2 };

constexpr int BM = 128;  // resident rows per CTA
constexpr int BN = 64;   // streamed rows per tile (= one SWIZZLE_128B row of bf16 in the T tile)

struct alignas(64) AParams {
  CUtensorMap tmR[2];
  CUtensorMap tmS[2];
  int Lr, Ls, H, D, B;
  float scale;
  bf16* out0;
  bf16* out1;
  long long o_rs, o_bs;
  float* lse;          // FWD: written; DQ/DKV: read.  [B, H, Lq]
  const float* delta;  // DQ/DKV.                      [B, H, Lq]
#ifdef PT_ATTN_TRACE
  unsigned long long* trace;   // development builds only (tools/attn_trace.py): [12 warps][1024 events][id, clock] of CTA 0
  int trace_last;
#endif
};

#ifdef PT_ATTN_TRACE
unsigned long long* g_attn_trace = nullptr;
int g_attn_trace_mode = -1;   // trace only kernels of this MODE (-1: all)
int g_attn_trace_last = 0;
// trace_last = 0: event log of CTA 0.  trace_last = 1: only the LAST event of every warp of every CTA, [CTA][12 warps][id | count << 32]:
// pointed at pinned host memory it survives a trapped kernel and shows where each warp of a dead pipeline is waiting.
#define TR(id)                                                                   \
  do {                                                                           \
    if (p.trace != nullptr && lane == 0) {                                       \
      if (p.trace_last) {                                                        \
        *(volatile unsigned long long*)(p.trace + blockIdx.x * 12 + warp) = (unsigned long long)(id) | ((unsigned long long)tr_n << 32); \
        ++tr_n;                                                                  \
      } else if (blockIdx.x == 0 && tr_n < 1024) {                               \
        p.trace[(warp * 1024 + tr_n) * 2] = (unsigned long long)(id);            \
        p.trace[(warp * 1024 + tr_n) * 2 + 1] = (unsigned long long)clock64();   \
        ++tr_n;                                                                  \
      }                                                                          \
    }                                                                            \
  } while (0)
#define PT_ATTN_SET_TRACE(ap) ((ap).trace = g_attn_trace, (ap).trace_last = g_attn_trace_last)
#else
#define TR(id)
#define PT_ATTN_SET_TRACE(ap)
#endif

template <int MODE, int DP>
struct ACfg {
  static constexpr int NX = MODE == MODE_FWD ? 1 : 2;       // stage-1 products per tile
  static constexpr int NACC = MODE == MODE_DKV ? 2 : 1;     // stage-2 accumulators written per tile
  static constexpr int NACC_T = MODE == MODE_FWD ? 2 : NACC; // accumulators held in TMEM (FWD: one per transform group, merged in the epilogue)
  static constexpr int XSLOTS = MODE == MODE_FWD ? (DP > 128 ? 1 : 2) : 1;   // FWD: X slots per group (TMEM permitting), so stage 1 runs a tile ahead
  // one X buffer / one transform group: DKV at DP = 192 for TMEM (2*128 + 2*192 > 512); DQ at DP = 192 so that shared memory
  // holds two ring stages (a single stage serialises the TMA round trip with every tile)
  static constexpr int XBUF = (MODE != MODE_FWD && DP > 128) ? 1 : 2;
  static constexpr int KB = DP / 64;                         // 64-element blocks along the head dimension
  static constexpr int R_BYTES = BM * DP * 2;
  static constexpr int S_BYTES = BN * DP * 2;
  static constexpr int STAGE_BYTES = 2 * S_BYTES;
  static constexpr int FIXED = NX * R_BYTES;                 // the transformed tiles (P, dS) live in TMEM, over the logits they came from
  static constexpr int BUDGET = (MODE == MODE_DKV ? 205 : 221) * 1024;                  // of 227 KB: leaves room for alignment slack, barriers, statistics
  static constexpr int NSTAGE_FIT = (BUDGET - FIXED) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_FIT >= 8 ? 8 : NSTAGE_FIT;   // deep ring: bytes in flight must cover the TMA round trip
  static constexpr int OFF_S = NX * R_BYTES;
  static constexpr int OFF_BAR = OFF_S + NSTAGE * STAGE_BYTES;
  static constexpr int OFF_STAT = OFF_BAR + 512;
  static constexpr int STAT_COLS = 2048;                     // DKV: query rows whose statistics fit in shared memory
  static constexpr int STAT_BYTES = (MODE == MODE_DKV ? 2 * STAT_COLS * 4 : 0) + 4 * BM * 4;   // DKV column stats [2][STAT_COLS] + row reduce [2][BM]
  static constexpr int SMEM_BYTES = 1024 + OFF_STAT + STAT_BYTES;
  static constexpr int XW = 128;                             // TMEM columns per X buffer (two 64-column products)
  static constexpr int X_COLS = MODE == MODE_FWD ? 2 * XSLOTS * BN : XBUF * XW;
  static constexpr int ACC_STRIDE = DP;
  // Backward: the transformed tiles (dS; P^T and dS^T) get their OWN TMEM columns when there is room (32 columns per 64 x bf16), so
  // the logit buffer is free as soon as the transform has READ it and stage 1 refills it while the transform still computes.  Written
  // over the logits (the first version) the buffer came back only after stage 2 had consumed them: transform -> stage 2 -> stage 1 ->
  // transform is ~1000 clocks (per-warp event trace, tools/attn_trace.py) during which the group had nothing to do.
  static constexpr int T_PER_G = MODE == MODE_FWD ? 0 : NACC * (BN / 2);
  static constexpr bool SEP_T = MODE != MODE_FWD && (XBUF * XW + XBUF * T_PER_G + NACC_T * ACC_STRIDE <= 512);
  static constexpr int T_COLS = SEP_T ? XBUF * T_PER_G : 0;
  static constexpr int TMEM_USED = X_COLS + T_COLS + NACC_T * ACC_STRIDE;
  static constexpr int TMEM_COLS = TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512));
  static_assert(NSTAGE >= 1, "smem budget");
  static_assert(TMEM_USED <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// FWD, 32 columns: P = exp2(x * c2 - mc) -> bf16 pairs -> 16 TMEM columns of this thread's lane (A operand of stage 2)
template <bool PARTIAL>
__device__ __forceinline__ void fwd_chunk(const uint32_t* v, int c, int ncol, float c2, float mc, float& rsum, uint32_t tdst) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float e0 = ex2f(fmaf(__uint_as_float(v[i]), c2, -mc));
    float e1 = ex2f(fmaf(__uint_as_float(v[i + 1]), c2, -mc));
    if (PARTIAL && c * 32 + i >= ncol) e0 = 0.f;
    if (PARTIAL && c * 32 + i + 1 >= ncol) e1 = 0.f;
    rsum += e0 + e1;
    pk[i >> 1] = pack2(e0, e1);
  }
  tmem_st16(tdst + c * 16, pk);
}

template <int MODE, int DP>
__global__ void __launch_bounds__(352, 1) attn_kernel(const __grid_constant__ AParams p) {
  using C = ACfg<MODE, DP>;
  constexpr int NX = C::NX, NACC = C::NACC, XBUF = C::XBUF, KB = C::KB, NSTAGE = C::NSTAGE, XSLOTS = C::XSLOTS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sR = sbase, sS = sbase + C::OFF_S, bar0 = sbase + C::OFF_BAR;
  // barriers (every one of them is reused across work items with a running phase)
  const uint32_t r_full = bar0;
  auto s_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto s_empty = [&](int s) { return bar0 + 8u * (9 + s); };
  auto x_full = [&](int g, int slot) { return bar0 + 8u * (17 + g * 2 + slot); };
  auto x_empty = [&](int g, int slot) { return bar0 + 8u * (21 + g * 2 + slot); };
  // P / dS of group g, X slot `slot`, are in TMEM.  One barrier PER SLOT: with two X slots the warps of a group may be a tile apart
  // (a warp that takes the rescale path waits on p_empty and does a TMEM round trip while the other three go on to the next tile,
  // whose logits are already in the other slot); on a single per-group barrier the fast warps' arrivals for tile k+1 completed the
  // phase of tile k, and stage 2 consumed P rows the slow warp had not written yet (round-1 `full_d40_self` NaN: always the 32 rows
  // of one warp, only with inputs that trigger rescales).  Per slot, a warp cannot arrive for tile k+2 before stage 2 of tile k.
  auto t_full = [&](int g, int slot) { return bar0 + 8u * (25 + g * 2 + slot); };
  auto p_empty = [&](int g, int slot) { return bar0 + 8u * (29 + g * 2 + slot); };  // stage 2 has consumed them: the X slot may be refilled
  const uint32_t acc_full = bar0 + 8u * 33;
  const uint32_t acc_empty = bar0 + 8u * 34;
  const uint32_t r_empty = bar0 + 8u * 35;
  constexpr int TMEM_SLOT_IDX = 36;
  const uint32_t tmem_slot = bar0 + 8u * TMEM_SLOT_IDX;
  float* sstat = reinterpret_cast<float*>(sgen + C::OFF_STAT);          // DKV: [2][STAT_COLS] column statistics of the work item
  float* sred = sstat + (MODE == MODE_DKV ? 2 * C::STAT_COLS : 0);      // [2][BM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef PT_ATTN_TRACE
  int tr_n = 0;
#endif
  const int n_tiles = (p.Ls + BN - 1) / BN;
  const int ks1 = (p.D + 15) >> 4;            // stage-1 k-steps (head dim, zero padded to a multiple of 16)
  const int nd = ks1 << 4;                    // stage-2 N
  // persistent CTA: work item w -> (resident row block, head, batch), row block fastest (neighbouring CTAs share K/V in L2)
  const int n_rblk = (p.Lr + BM - 1) / BM;
  const int total_work = n_rblk * p.H * p.B;
  auto decode = [&](int w, int& r0, int& h, int& b) {
    r0 = (w % n_rblk) * BM;
    w /= n_rblk;
    h = w % p.H;
    b = w / p.H;
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NX; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmR[i])) : "memory");
    for (int i = 0; i < 2; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmS[i])) : "memory");
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(s_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      for (int u = 0; u < 2; ++u) {
        mbar_init(x_full(s, u), 1);
        mbar_init(x_empty(s, u), 128);
        mbar_init(p_empty(s, u), 1);
        mbar_init(t_full(s, u), 128);
      }
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(sgen + C::OFF_BAR + 8 * TMEM_SLOT_IDX);
  // X buffers: FWD gives each group two 64-column slots (stage 1 runs a whole tile ahead of the group); the backward
  // modes need X1 | X2 per tile and have TMEM for one 128-column buffer per group only
  auto xcol = [&](int g, int slot_or_x) { return (uint32_t)(MODE == MODE_FWD ? (g * XSLOTS + slot_or_x) * BN : g * C::XW + slot_or_x * BN); };
  auto acccol = [&](int a) { return (uint32_t)(C::X_COLS + C::T_COLS + a * C::ACC_STRIDE); };
  // backward: where transformed tile `a` of group g goes (its own columns, or over logit product `a`)
  auto tcol = [&](int g, int a) { return C::SEP_T ? (uint32_t)(C::X_COLS + g * C::T_PER_G + a * (BN / 2)) : xcol(g, a); };
  auto gsel = [&](int j) { return XBUF == 2 ? (j & 1) : 0; };
  // tiles of one sweep handled by group g
  const int per_g0 = XBUF == 2 ? (n_tiles + 1) >> 1 : n_tiles, per_g1 = XBUF == 2 ? n_tiles >> 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (warp-uniform control flow, one elected lane issues)
    const bool leader = elect_one();
    int s = 0, ph = 1;   // ring position / parity to wait for on s_empty
    int wi = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wi) {
      int r0, h, b;
      decode(w, r0, h, b);
      mbar_wait(r_empty, (wi & 1) ^ 1);   // stage 1 of the previous work item has finished reading the resident tiles
      if (leader) {
        mbar_expect_tx(r_full, NX * C::R_BYTES);
#pragma unroll
        for (int x = 0; x < NX; ++x)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) tma_load_4d(sR + x * C::R_BYTES + kb * (BM * 128), &p.tmR[x], r_full, kb * 64, r0, h, b);
      }
      auto load_tile = [&](const CUtensorMap* t0, const CUtensorMap* t1, int row0, int row1) {
        TR(40);
        mbar_wait(s_empty(s), ph);
        TR(41);
        if (leader) {
          mbar_expect_tx(s_full(s), 2 * C::S_BYTES);
          const uint32_t dst = sS + s * C::STAGE_BYTES;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) tma_load_4d(dst + kb * (BN * 128), t0, s_full(s), kb * 64, row0, h, b);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) tma_load_4d(dst + C::S_BYTES + kb * (BN * 128), t1, s_full(s), kb * 64, row1, h, b);
        }
        if (++s == NSTAGE) s = 0, ph ^= 1;
      };
      for (int j = 0; j < n_tiles; ++j) load_tile(&p.tmS[0], &p.tmS[1], j * BN, j * BN);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ stage-1 MMA issuer (warp-uniform control flow, one elected lane issues).
    // Runs ahead of the transform groups as far as the ring and the X buffers allow; never waits on stage 2.  Per tile it waits for the
    // ring stage and the X buffer, converts ~25 operands to uniform registers and issues up to 8 MMAs + a commit: 1000-1400 clocks of
    // dependent scalar work per tile (per-warp event trace, tools/attn_trace.py).  A second issuer warp (one per group) was tried:
    // 3 % faster forward, but the d = 160 forward died with a launch failure within seconds (not a barrier time-out: no wait had
    // spun long) in every build that had two warps issuing stage-1 MMAs, and never with one -- not shipped.
    constexpr uint32_t gmask = 3u;      // groups served by this issuer (bit g)
    if (gmask != 0) {
      const bool leader = elect_one();
      const uint32_t idesc1 = idesc_f16(0, 0, BN, BM);
      // UMMA descriptors are linear in the shared-memory address: precompute the bases, add (bytes >> 4) per use
      const uint64_t dR = umma_desc(sR, 0, 1024);            // resident tiles, K-major
      const uint64_t dS = umma_desc(sS, 0, 1024);            // streamed tiles, K-major view
      int tbase = 0;         // streamed tiles of earlier work items (ring position of tile j = (tbase + j) % NSTAGE)
      int kb0 = 0, kb1 = 0;  // X fills by earlier work items, per group
      // FWD: an X slot is released by the stage-2 commit (P was written over the logits and has been consumed).  Backward with separate
      // T columns: by the transform threads once they have read the logits (x_empty).  Per slot i = g * 2 + slot: what the last fill was
      // (2 bits: 0 none, 1 released by x_empty, 2 by p_empty) and the parity of the next phase of each of its two barriers.
      uint32_t last_kind = 0, par_x = 0, par_p = 0;
      auto wait_slot_free = [&](int g, int slot, uint32_t kind) {
        const int i = g * 2 + slot;
        const uint32_t prev = (last_kind >> (2 * i)) & 3u;
        if (prev == 1u) {
          mbar_wait(x_empty(g, slot), (par_x >> i) & 1u);
          par_x ^= 1u << i;
        } else if (prev == 2u) {
          mbar_wait(p_empty(g, slot), (par_p >> i) & 1u);
          par_p ^= 1u << i;
        }
        last_kind = (last_kind & ~(3u << (2 * i))) | (kind << (2 * i));
      };
      int wi = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wi) {
        mbar_wait(r_full, wi & 1);
        for (int j = 0; j < n_tiles; ++j) {
          const int g = gsel(j);
          if (!((gmask >> g) & 1u)) continue;
          const int t = tbase + j;
          const int s1 = t % NSTAGE;
          const uint32_t ph1 = (uint32_t)(t / NSTAGE) & 1u;
          // FWD: the group's k-th tile overall uses X slot k % XSLOTS; backward: one buffer (X1 | X2)
          const int k = (g ? kb1 : kb0) + (XBUF == 2 ? (j >> 1) : j);
          const int slot = (MODE == MODE_FWD && XSLOTS == 2) ? (k & 1) : 0;
          TR(20);
          mbar_wait(s_full(s1), ph1);
          TR(21);
          wait_slot_free(g, slot, (MODE != MODE_FWD && C::SEP_T) ? 1u : 2u);
          TR(22);
          fence_after();
          if (leader) {
            const uint64_t bS = dS + (uint64_t)(s1 * (C::STAGE_BYTES >> 4));
#pragma unroll
            for (int x = 0; x < NX; ++x) {
              const uint64_t a0 = dR + (uint64_t)(x * (C::R_BYTES >> 4));
              const uint64_t b0 = bS + (uint64_t)(x * (C::S_BYTES >> 4));
              const uint32_t dcol = tmem + xcol(g, MODE == MODE_FWD ? slot : x);
#pragma unroll
              for (int kk = 0; kk < DP / 16; ++kk)
                if (kk < ks1)
                  umma_f16(dcol, a0 + (uint64_t)((kk >> 2) * (BM * 128 >> 4) + (kk & 3) * 2),
                           b0 + (uint64_t)((kk >> 2) * (BN * 128 >> 4) + (kk & 3) * 2), idesc1, kk > 0 ? 1u : 0u);
            }
            umma_commit(x_full(g, slot));
          }
          __syncwarp();
          TR(23);
        }
        if (leader) umma_commit(r_empty);      // the stage-1 MMAs of the work item have read the resident tiles
        __syncwarp();
        tbase += n_tiles;
        kb0 += per_g0;
        kb1 += per_g1;
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ stage-2 MMA issuer: ACC += P * S, A = P from TMEM (written by the
    // transform over the logits), B = MN-major view of the streamed tile; frees the ring slot and the X slot
    const bool leader = elect_one();
    const uint32_t idesc2 = idesc_f16(0, 1, nd, BM);
    const uint64_t dSmn = umma_desc(sS, BN * 128, 1024);   // streamed tiles, MN-major view
    int s2 = 0;
    int tu0 = 0, tu1 = 0;   // t_full phases consumed so far, per group
    int wi = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wi) {
      TR(28);
      mbar_wait(acc_empty, (wi & 1) ^ 1);                    // the epilogue of the previous work item has drained the accumulators
      TR(29);
      fence_after();
      for (int j = 0; j < n_tiles; ++j) {
        const int g = gsel(j);
        const int use = (g ? tu1 : tu0) + (XBUF == 2 ? (j >> 1) : j);      // fills of this group's X so far (== kf for FWD)
        const int slot = (MODE == MODE_FWD && XSLOTS == 2) ? (use & 1) : 0;
        TR(30);
        mbar_wait(t_full(g, slot), ((MODE == MODE_FWD && XSLOTS == 2) ? (use >> 1) : use) & 1);
        TR(31);
        fence_after();
        if (leader) {
          const uint64_t bS = dSmn + (uint64_t)(s2 * (C::STAGE_BYTES >> 4));
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            // B operand: FWD -> V (S2); DQ -> K (S1); DKV: dV <- dO (S2), dK <- Q (S1)
            const int cs = MODE == MODE_FWD ? 1 : (MODE == MODE_DQ ? 0 : (a == 0 ? 1 : 0));
            // A operand in TMEM: FWD P in its X slot; DQ dS over X1; DKV P^T over X1 (-> dV), dS^T over X2 (-> dK)
            const uint32_t acol = tmem + (MODE == MODE_FWD ? xcol(g, slot) : tcol(g, a));
            const uint64_t b0 = bS + (uint64_t)(cs * (C::S_BYTES >> 4));
            const uint32_t dcol = tmem + acccol(MODE == MODE_FWD ? g : a);     // FWD: each group accumulates into its own O
            const bool first = MODE == MODE_FWD ? j < 2 : j == 0;
#pragma unroll
            for (int k = 0; k < BN / 16; ++k)
              umma_f16_ts(dcol, acol + (uint32_t)(k * 8), b0 + (uint64_t)(k * (2048 >> 4)), idesc2, (first && k == 0) ? 0u : 1u);
          }
          umma_commit(p_empty(g, slot));
          umma_commit(s_empty(s2));
          if (j == n_tiles - 1) umma_commit(acc_full);
        }
        __syncwarp();
        TR(33);
        if (++s2 == NSTAGE) s2 = 0;
      }
      tu0 += per_g0;
      tu1 += per_g1;
    }
  } else {
    // ------------------------------------------------------------------ transform groups (warps 2-5, 6-9)
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;                // row of the 128-row tile owned by this thread
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const bool active = XBUF == 2 || g == 0;      // with one X buffer only group 0 transforms (group 1 helps in the epilogue)
    const int jstep = XBUF == 2 ? 2 : 1;
    const float c2 = p.scale * 1.4426950408889634f;
    // FWD, exponent turn-taking.  Warp q of group 0 and warp q of group 1 share a scheduler and its MUFU unit (16 ex2/clk/SM).  Left
    // alone the two drift into the SAME phase: both exponentiate at half rate (~1100 clocks), then both sit in the latency-bound
    // part (row maximum, TMEM round trips, barriers) with the MUFU idle (per-warp event trace: MUFU 49 % busy).  A token per warp
    // pair forces the exponent phases to alternate, so one warp's latencies hide under the other's exponentials.  Two named
    // barriers per pair: the owner waits with bar.sync (64 = its 32 threads + the partner's 32 arrivals), the partner hands the turn
    // over with bar.arrive after its own phase.  Turns follow the tile order (tile j belongs to group j & 1); with an odd number of
    // tiles group 1 passes once without work so that the next work item starts with group 0 again.
#ifdef PT_ATTN_NO_TURNS
    constexpr bool TURNS = false;
#else
    constexpr bool TURNS = MODE == MODE_FWD && XBUF == 2;
#endif
    const int bar_mine = 3 + 2 * q + g, bar_other = 3 + 2 * q + (g ^ 1);
    auto turn_wait = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_mine) : "memory"); };
    auto turn_pass = [&]() { asm volatile("bar.arrive %0, 64;" ::"r"(bar_other) : "memory"); };
    if (TURNS && g == 1) turn_pass();      // group 0 moves first
    int kx = 0;           // running X use count of this group (across work items)
    uint32_t v1[32], v2[32];
    // Backward: the softmax statistics of a work item (lse, delta: global loads, a DRAM round trip) are requested one work item ahead --
    // before the epilogue of the previous one -- and only consumed here; loaded on demand they stalled every work item for the full
    // latency (per-warp event trace: 1300 clocks in dQ, 2300 in dK/dV per work item).
    constexpr int NPRE = 4;                 // DKV: statistics columns per thread held in registers (4 x 256 = 1024 query rows; beyond: on demand)
    float pre_a[MODE == MODE_DKV ? NPRE : 1], pre_d[MODE == MODE_DKV ? NPRE : 1];
    auto prefetch_stats = [&](int w2) {
      if (MODE == MODE_FWD) return;
      int r0n, hn, bn_;
      decode(w2, r0n, hn, bn_);
      if (MODE == MODE_DQ) {
        pre_a[0] = INFINITY, pre_d[0] = 0.f;
        if (w2 < total_work && r0n + row < p.Lr) {
          const long long si = ((long long)bn_ * p.H + hn) * p.Lr + r0n + row;
          pre_a[0] = __ldg(p.lse + si);
          pre_d[0] = __ldg(p.delta + si);
        }
      } else {
        const long long sb = ((long long)bn_ * p.H + hn) * p.Ls;
#pragma unroll
        for (int u = 0; u < NPRE; ++u) {
          const int qi = (int)threadIdx.x - 64 + u * 256;
          pre_a[u] = INFINITY, pre_d[u] = 0.f;
          if (w2 < total_work && qi < p.Ls) {
            pre_a[u] = __ldg(p.lse + sb + qi);
            pre_d[u] = __ldg(p.delta + sb + qi);
          }
        }
      }
    };
    prefetch_stats(blockIdx.x);
    int wi = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++wi) {
      int r0, h, b;
      decode(w, r0, h, b);
      const int r = r0 + row;

      if (MODE == MODE_FWD) {
        // ---- online softmax: this group keeps its own running maximum m (raw logit units), row sum l and accumulator O_g.
        // The reference point m only moves when a row's maximum grows by more than 2^8 (then O_g and l are rescaled), so
        // p = exp2((x - m) c) <= 256 and the rescale is rare; the two groups are merged exactly in the epilogue.
        float m = -INFINITY, l = 0.f;
        int kw = 0;                                      // tiles of this group in this work item so far
        for (int j = g; j < n_tiles; j += 2, ++kx, ++kw) {
          const int slot = XSLOTS == 2 ? (kx & 1) : 0;
          TR(0);
          mbar_wait(x_full(g, slot), (XSLOTS == 2 ? (kx >> 1) : kx) & 1);
          TR(1);
          fence_after();
          tmem_ld32(tl + xcol(g, slot), v1);
          tmem_ld32(tl + xcol(g, slot) + 32, v2);
          tmem_wait_ld();
          TR(2);
          const int ncol = min(BN, p.Ls - j * BN);
          float mt = -INFINITY;
          if (ncol == BN) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // four independent chains (a single one is 32 dependent FMNMX3)
#pragma unroll
            for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(v1[i]), __uint_as_float(v2[i])));
            mt = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < ncol) mt = fmaxf(mt, __uint_as_float(v1[i]));
              if (32 + i < ncol) mt = fmaxf(mt, __uint_as_float(v2[i]));
            }
          }
          const float m_new = fmaxf(m, mt);
          if (kw == 0) {
            m = m_new;                                   // nothing accumulated yet (stage 2 of this tile overwrites O_g)
          } else {
            const bool grow = (m_new - m) * c2 > 8.f;
            if (__any_sync(0xffffffffu, grow)) {         // warp-collective: tcgen05.ld / st need the whole warp
              const float alpha = grow ? ex2f((m - m_new) * c2) : 1.f;
              if (grow) m = m_new;
              l *= alpha;
              // stage 2 of this group's previous tile must have retired before O_g is touched
              const int kp = kx - 1;
              mbar_wait(p_empty(g, XSLOTS == 2 ? (kp & 1) : 0), (XSLOTS == 2 ? (kp >> 1) : kp) & 1);
              fence_after();
              for (int cc = 0; cc * 16 < p.D; ++cc) {
                uint32_t t[16];
                tmem_ld16(tl + acccol(g) + cc * 16, t);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
                tmem_st16(tl + acccol(g) + cc * 16, t);
              }
              tmem_wait_st();
            }
          }
          TR(3);
          const float mc = m * c2;
          const uint32_t tdst = tl + xcol(g, slot);      // P overwrites the logits of this slot (32 of its 64 columns)
          if (TURNS) turn_wait();
          if (ncol == BN) {
            fwd_chunk<false>(v1, 0, ncol, c2, mc, l, tdst);
            fwd_chunk<false>(v2, 1, ncol, c2, mc, l, tdst);
          } else {
            fwd_chunk<true>(v1, 0, ncol, c2, mc, l, tdst);
            fwd_chunk<true>(v2, 1, ncol, c2, mc, l, tdst);
          }
          if (TURNS) turn_pass();
          TR(4);
          tmem_wait_st();
          TR(5);
          fence_before();
          mbar_arrive(t_full(g, slot));
          TR(6);
        }
        if (TURNS && g == 1 && (n_tiles & 1)) {      // odd tile count: the last turn was group 0's and so is the next one
          turn_wait();
          turn_pass();
        }
        TR(10);
        // ---- merge the two groups: M = max(m_0, m_1), w_g = exp2((m_g - M) c), L = sum w_g l_g, O = sum w_g O_g / L
        sred[(g * BM + row) * 2] = m;
        sred[(g * BM + row) * 2 + 1] = l;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float m0 = sred[row * 2], l0 = sred[row * 2 + 1], m1 = sred[(BM + row) * 2], l1 = sred[(BM + row) * 2 + 1];
        const float mm = fmaxf(m0, m1);
        const float w0 = ex2f((m0 - mm) * c2), w1 = n_tiles > 1 ? ex2f((m1 - mm) * c2) : 0.f;    // a group without tiles has m = -inf, l = 0
        const float lsum = w0 * l0 + w1 * l1;
        TR(11);
        mbar_wait(acc_full, wi & 1);
        TR(12);
        fence_after();
        const float f0 = w0 / lsum, f1 = w1 / lsum;
        bf16* orow = p.out0 + (long long)b * p.o_bs + (long long)r * p.o_rs + (long long)h * p.D;
        TR(15);
        for (int cc = g; cc * 16 < p.D; cc += 2) {
          tmem_ld16(tl + acccol(0) + cc * 16, v1);
          if (n_tiles > 1) tmem_ld16(tl + acccol(1) + cc * 16, v2);
          tmem_wait_ld();
          TR(16);
          if (r < p.Lr) {
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(v1[i]) * f0;
            if (n_tiles > 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = fmaf(__uint_as_float(v2[i]), f1, o[i]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
              if (cc * 16 + u * 8 < p.D) {
                uint4 wv;
                wv.x = pack2(o[u * 8], o[u * 8 + 1]);
                wv.y = pack2(o[u * 8 + 2], o[u * 8 + 3]);
                wv.z = pack2(o[u * 8 + 4], o[u * 8 + 5]);
                wv.w = pack2(o[u * 8 + 6], o[u * 8 + 7]);
                *reinterpret_cast<uint4*>(orow + cc * 16 + u * 8) = wv;
              }
          }
          TR(17);
        }
        fence_before();
        mbar_arrive(acc_empty);
        TR(18);
        if (g == 0 && r < p.Lr) p.lse[((long long)b * p.H + h) * p.Lr + r] = mm * p.scale + __logf(lsum);
        TR(13);
        asm volatile("bar.sync 2, 256;" ::: "memory");   // sred is rewritten by the next work item
        TR(14);
      } else {
        // ---- backward modes
        // DQ: this thread's row statistics (+inf masks a row beyond Lr)
        const float lse2 = MODE == MODE_DQ ? pre_a[0] * 1.4426950408889634f : INFINITY, dl = MODE == MODE_DQ ? pre_d[0] * p.scale : 0.f;
        if (MODE == MODE_DKV) {
          // column statistics of the whole work item (columns = query rows), once: lse * log2(e) (+inf masks a column), delta * scale
          asm volatile("bar.sync 2, 256;" ::: "memory");     // everyone has finished reading the previous work item's statistics
          const long long sb = ((long long)b * p.H + h) * p.Ls;
#pragma unroll
          for (int u = 0; u < NPRE; ++u) {
            const int qi = (int)threadIdx.x - 64 + u * 256;
            if (qi < n_tiles * BN) {
              sstat[qi] = pre_a[u] * 1.4426950408889634f;
              sstat[C::STAT_COLS + qi] = pre_d[u] * p.scale;
            }
          }
          for (int qi = threadIdx.x - 64 + NPRE * 256; qi < n_tiles * BN; qi += 256) {
            float a = INFINITY, d = 0.f;
            if (qi < p.Ls) {
              a = __ldg(p.lse + sb + qi) * 1.4426950408889634f;
              d = __ldg(p.delta + sb + qi) * p.scale;
            }
            sstat[qi] = a;
            sstat[C::STAT_COLS + qi] = d;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        if (active)
          for (int j = g; j < n_tiles; j += jstep, ++kx) {
            const float4* cst = reinterpret_cast<const float4*>(sstat + j * BN);
            const float4* cdl = reinterpret_cast<const float4*>(sstat + C::STAT_COLS + j * BN);
            TR(0);
            mbar_wait(x_full(g, 0), kx & 1);
            TR(1);
            fence_after();
            // 32 logit columns of both products -> P / dS as bf16 pairs
            auto chunk = [&](int c, const uint32_t* x1, const uint32_t* x2, uint32_t* pp, uint32_t* dd) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float l4[4] = {lse2, lse2, lse2, lse2}, d4[4] = {dl, dl, dl, dl};
                if (MODE == MODE_DKV) {
                  const float4 a = cst[(c * 32 + i) >> 2], d = cdl[(c * 32 + i) >> 2];
                  l4[0] = a.x, l4[1] = a.y, l4[2] = a.z, l4[3] = a.w;
                  d4[0] = d.x, d4[1] = d.y, d4[2] = d.z, d4[3] = d.w;
                }
                float pe[4], de[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float e = ex2f(fmaf(__uint_as_float(x1[i + u]), c2, -l4[u]));
                  pe[u] = e;
                  de[u] = e * fmaf(__uint_as_float(x2[i + u]), p.scale, -d4[u]);
                }
                pp[i >> 1] = pack2(pe[0], pe[1]);
                pp[(i >> 1) + 1] = pack2(pe[2], pe[3]);
                dd[i >> 1] = pack2(de[0], de[1]);
                dd[(i >> 1) + 1] = pack2(de[2], de[3]);
              }
            };
            auto wait_t_free = [&]() {                  // stage 2 of this group's previous tile has consumed the T columns
              if (C::SEP_T && kx > 0) {
                mbar_wait(p_empty(g, 0), (kx - 1) & 1);
                fence_after();
              }
            };
            {
#pragma unroll
              for (int c = 0; c < BN / 32; ++c) {
                tmem_ld32(tl + xcol(g, 0) + c * 32, v1);
                tmem_ld32(tl + xcol(g, 1) + c * 32, v2);
                tmem_wait_ld();
                if (C::SEP_T && c == BN / 32 - 1) {       // the logits of this tile are in registers: stage 1 may refill the buffer
                  fence_before();
                  mbar_arrive(x_empty(g, 0));
                }
                uint32_t pp[16], dd[16];
                chunk(c, v1, v2, pp, dd);
                if (c == 0) wait_t_free();
                // DQ: dS; DKV: P^T and dS^T (16 columns per 32 logits) -- in their own columns, or in place over the logits already consumed
                if (MODE == MODE_DKV) {
                  tmem_st16(tl + tcol(g, 0) + c * 16, pp);
                  tmem_st16(tl + tcol(g, 1) + c * 16, dd);
                } else {
                  tmem_st16(tl + tcol(g, 0) + c * 16, dd);
                }
              }
            }
            TR(4);
            tmem_wait_st();
            TR(5);
            fence_before();
            mbar_arrive(t_full(g, 0));
            TR(6);
          }
        prefetch_stats(w + (int)gridDim.x);      // the next work item's statistics: in flight during the epilogue
        // ---- epilogue: accumulators -> bf16 rows
        TR(11);
        mbar_wait(acc_full, wi & 1);
        TR(12);
        fence_after();
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          bf16* obase = (a == 0 ? p.out0 : p.out1) + (long long)b * p.o_bs + (long long)r * p.o_rs + (long long)h * p.D;
          // DKV: group a writes accumulator a; DQ: the two groups interleave 16-column chunks
          const int c_begin = MODE == MODE_DKV ? 0 : g, c_step = MODE == MODE_DKV ? 1 : 2;
          if (MODE == MODE_DKV && g != a) continue;
          for (int cc = c_begin; cc * 16 < p.D; cc += c_step) {
            tmem_ld16(tl + acccol(a) + cc * 16, v1);
            tmem_wait_ld();
            if (r < p.Lr) {
#pragma unroll
              for (int u = 0; u < 2; ++u)
                if (cc * 16 + u * 8 < p.D) {
                  uint4 wv;
                  wv.x = pack2(__uint_as_float(v1[u * 8]), __uint_as_float(v1[u * 8 + 1]));
                  wv.y = pack2(__uint_as_float(v1[u * 8 + 2]), __uint_as_float(v1[u * 8 + 3]));
                  wv.z = pack2(__uint_as_float(v1[u * 8 + 4]), __uint_as_float(v1[u * 8 + 5]));
                  wv.w = pack2(__uint_as_float(v1[u * 8 + 6]), __uint_as_float(v1[u * 8 + 7]));
                  *reinterpret_cast<uint4*>(obase + cc * 16 + u * 8) = wv;
                }
            }
          }
        }
        fence_before();
        mbar_arrive(acc_empty);
        TR(13);
      }
    }
    if (TURNS && g == 0) turn_wait();      // consume group 1's last hand-over: the named barriers are left clean
  }

  fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
  }
}

// delta[b, h, q] = sum_j dO[b, q, h*d + j] * O[b, q, h*d + j].  A CTA takes DELTA_ROWS rows; consecutive threads read consecutive
// 16-byte vectors of a row (coalesced), each vector lies inside one head (d % 8 == 0), partial dots meet in shared memory.
constexpr int DELTA_ROWS = 8;
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ o, long long o_rs, long long o_bs, const bf16* __restrict__ d_o,
                                                         long long do_rs, long long do_bs, float* __restrict__ delta, int B, int H, int L, int D) {
  __shared__ float acc[DELTA_ROWS * 64];   // [row][head], H <= 64
  const long long row0 = (long long)blockIdx.x * DELTA_ROWS;
  const long long nrows = (long long)B * L;
  const int nvec = (H * D) >> 3, vph = D >> 3;
  for (int i = threadIdx.x; i < DELTA_ROWS * H; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < DELTA_ROWS * nvec; i += blockDim.x) {
    const int rl = i / nvec, v = i - rl * nvec;
    const long long row = row0 + rl;
    if (row < nrows) {
      const int bb = (int)(row / L), l = (int)(row - (long long)bb * L);
      float a[8], c[8];
      load8(o + (long long)bb * o_bs + (long long)l * o_rs + v * 8, a);
      load8(d_o + (long long)bb * do_bs + (long long)l * do_rs + v * 8, c);
      float t = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) t = fmaf(a[u], c[u], t);
      atomicAdd(&acc[rl * H + v / vph], t);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DELTA_ROWS * H; i += blockDim.x) {
    const int rl = i / H, hh = i - rl * H;
    const long long row = row0 + rl;
    if (row < nrows) {
      const int bb = (int)(row / L), l = (int)(row - (long long)bb * L);
      delta[((long long)bb * H + hh) * L + l] = acc[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
int encode_heads(CUtensorMap* tm, const void* ptr, int D, int L, int H, int B, long long rs, long long bs, int box_rows, const char* name) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    pt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return PT_ECUDA;
  }
  PT_REQUIRE(ptr != nullptr, "pt_attn: %s is null", name);
  PT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && rs % 8 == 0 && bs % 8 == 0, "pt_attn: %s needs a 16-byte aligned base and strides that are multiples of 8", name);
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)rs * 2, (cuuint64_t)D * 2, (cuuint64_t)(B > 1 ? bs : 8) * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pt_set_error("pt_attn: cuTensorMapEncodeTiled(%s) failed: CUresult %d (D=%d L=%d H=%d B=%d rs=%lld bs=%lld)", name, (int)r, D, L, H, B, rs, bs);
    return PT_ECUDA;
  }
  return PT_OK;
}

template <int MODE, int DP>
int launch_attn(const AParams& ap, dim3 grid, cudaStream_t st) {
  PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(attn_kernel<MODE, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ACfg<MODE, DP>::SMEM_BYTES));
  const long long work = (long long)grid.x * grid.y * grid.z;   // (row blocks, heads, batch) -> one persistent CTA per SM
  const int sms = pt_num_sms();
#ifdef PT_ATTN_TRACE
  AParams apt = ap;
  if (g_attn_trace_mode >= 0 && g_attn_trace_mode != MODE) apt.trace = nullptr;
  attn_kernel<MODE, DP><<<(unsigned)(work < sms ? work : sms), 352, ACfg<MODE, DP>::SMEM_BYTES, st>>>(apt);
#else
  attn_kernel<MODE, DP><<<(unsigned)(work < sms ? work : sms), 352, ACfg<MODE, DP>::SMEM_BYTES, st>>>(ap);
#endif
  PT_LAUNCH_CHECK();
  return PT_OK;
}

template <int MODE>
int launch_mode(const AParams& ap, dim3 grid, cudaStream_t st) {
  if (ap.D <= 64) return launch_attn<MODE, 64>(ap, grid, st);
  if (ap.D <= 128) return launch_attn<MODE, 128>(ap, grid, st);
  return launch_attn<MODE, 192>(ap, grid, st);
}

int check_common(const pt_attn_t* a) {
  PT_REQUIRE(a != nullptr, "pt_attn: null descriptor");
  PT_REQUIRE(a->B >= 1 && a->H >= 1 && a->Lq >= 1 && a->Lk >= 1, "pt_attn: B=%d H=%d Lq=%d Lk=%d", a->B, a->H, a->Lq, a->Lk);
  PT_REQUIRE(a->d >= 8 && a->d <= 192 && a->d % 8 == 0, "pt_attn: head dim %d must be a multiple of 8 in [8, 192]", a->d);
  PT_REQUIRE(a->H <= 65535 && a->B <= 65535, "pt_attn: grid too large");
  PT_REQUIRE(a->lse != nullptr, "pt_attn: lse is null");
  return PT_OK;
}

}  // namespace

extern "C" int pt_attn_fwd(const pt_attn_t* a, void* stream) {
  if (int r = check_common(a)) return r;
  PT_REQUIRE(a->o != nullptr && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0 && a->o_rs % 8 == 0 && a->o_bs % 8 == 0, "pt_attn_fwd: output alignment");
  AParams ap;
  memset(&ap, 0, sizeof(ap));
    PT_ATTN_SET_TRACE(ap);
  if (int r = encode_heads(&ap.tmR[0], a->q, a->d, a->Lq, a->H, a->B, a->q_rs, a->q_bs, BM, "q")) return r;
  if (int r = encode_heads(&ap.tmS[0], a->k, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BN, "k")) return r;
  if (int r = encode_heads(&ap.tmS[1], a->v, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BN, "v")) return r;
  ap.Lr = a->Lq;
  ap.Ls = a->Lk;
  ap.H = a->H;
  ap.B = a->B;
  ap.D = a->d;
  ap.scale = a->scale;
  ap.out0 = reinterpret_cast<bf16*>(a->o);
  ap.o_rs = a->o_rs;
  ap.o_bs = a->o_bs;
  ap.lse = a->lse;
  dim3 grid((a->Lq + BM - 1) / BM, a->H, a->B);
  return launch_mode<MODE_FWD>(ap, grid, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int pt_attn_bwd(const pt_attn_t* a, void* stream) {
  if (int r = check_common(a)) return r;
  PT_REQUIRE(a->o && a->d_o && a->dq && a->dk && a->dv && a->delta, "pt_attn_bwd: null pointer");
  PT_REQUIRE(a->Lq <= 2048 - BN, "pt_attn_bwd: Lq=%d exceeds the %d query rows whose statistics are staged in shared memory", a->Lq, 2048 - BN);
  PT_REQUIRE((reinterpret_cast<uintptr_t>(a->o) & 15) == 0 && a->o_rs % 8 == 0 && a->o_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(a->dq) & 15) == 0 &&
                 a->dq_rs % 8 == 0 && a->dq_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(a->dk) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a->dv) & 15) == 0 && a->dkv_rs % 8 == 0 && a->dkv_bs % 8 == 0,
             "pt_attn_bwd: alignment");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    PT_REQUIRE(a->H <= 64, "pt_attn_bwd: at most 64 heads (H=%d)", a->H);
    const long long nrows = (long long)a->B * a->Lq;
    attn_delta_kernel<<<(unsigned)((nrows + DELTA_ROWS - 1) / DELTA_ROWS), 256, 0, st>>>(
        reinterpret_cast<const bf16*>(a->o), a->o_rs, a->o_bs, reinterpret_cast<const bf16*>(a->d_o), a->do_rs, a->do_bs, a->delta, a->B, a->H,
        a->Lq, a->d);
    PT_LAUNCH_CHECK();
  }
  {  // dQ: resident Q, dO ; streamed K, V
    AParams ap;
    memset(&ap, 0, sizeof(ap));
    PT_ATTN_SET_TRACE(ap);
    if (int r = encode_heads(&ap.tmR[0], a->q, a->d, a->Lq, a->H, a->B, a->q_rs, a->q_bs, BM, "q")) return r;
    if (int r = encode_heads(&ap.tmR[1], a->d_o, a->d, a->Lq, a->H, a->B, a->do_rs, a->do_bs, BM, "do")) return r;
    if (int r = encode_heads(&ap.tmS[0], a->k, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BN, "k")) return r;
    if (int r = encode_heads(&ap.tmS[1], a->v, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BN, "v")) return r;
    ap.Lr = a->Lq;
    ap.Ls = a->Lk;
    ap.H = a->H;
  ap.B = a->B;
    ap.D = a->d;
    ap.scale = a->scale;
    ap.out0 = reinterpret_cast<bf16*>(a->dq);
    ap.o_rs = a->dq_rs;
    ap.o_bs = a->dq_bs;
    ap.lse = a->lse;
    ap.delta = a->delta;
    dim3 grid((a->Lq + BM - 1) / BM, a->H, a->B);
    if (int r = launch_mode<MODE_DQ>(ap, grid, st)) return r;
  }
  {  // dK, dV: resident K, V ; streamed Q, dO
    AParams ap;
    memset(&ap, 0, sizeof(ap));
    PT_ATTN_SET_TRACE(ap);
    if (int r = encode_heads(&ap.tmR[0], a->k, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BM, "k")) return r;
    if (int r = encode_heads(&ap.tmR[1], a->v, a->d, a->Lk, a->H, a->B, a->kv_rs, a->kv_bs, BM, "v")) return r;
    if (int r = encode_heads(&ap.tmS[0], a->q, a->d, a->Lq, a->H, a->B, a->q_rs, a->q_bs, BN, "q")) return r;
    if (int r = encode_heads(&ap.tmS[1], a->d_o, a->d, a->Lq, a->H, a->B, a->do_rs, a->do_bs, BN, "do")) return r;
    ap.Lr = a->Lk;
    ap.Ls = a->Lq;
    ap.H = a->H;
  ap.B = a->B;
    ap.D = a->d;
    ap.scale = a->scale;
    ap.out0 = reinterpret_cast<bf16*>(a->dv);
    ap.out1 = reinterpret_cast<bf16*>(a->dk);
    ap.o_rs = a->dkv_rs;
    ap.o_bs = a->dkv_bs;
    ap.lse = a->lse;
    ap.delta = a->delta;
    dim3 grid((a->Lk + BM - 1) / BM, a->H, a->B);
    if (int r = launch_mode<MODE_DKV>(ap, grid, st)) return r;
  }
  return PT_OK;
}

#ifdef PT_ATTN_TRACE
// development builds only: per-warp event log of CTA 0 (see TR above); buf = [12][1024][2] uint64, zero-filled by the caller
extern "C" int pt_attn_set_trace(void* buf, int mode) {
  g_attn_trace_mode = (mode < 0 || mode == 7) ? -1 : (mode & 3);   // 7: every kernel, last-event form
  g_attn_trace_last = mode >= 0 && (mode & 4) ? 1 : 0;      // mode | 4: last-event-per-warp form (all CTAs)
  g_attn_trace = reinterpret_cast<unsigned long long*>(buf);
  unsigned long long* wp = (buf != nullptr && g_attn_trace_last) ? g_attn_trace + 160 * 12 : nullptr;   // after the last-event table
  cudaMemcpyToSymbol(g_watch, &wp, sizeof(wp));
  return PT_OK;
}
#endif
