// tcgen05 / TMEM / TMA GEMM family for sm_100a.
//
//   out[z, m, n] = alpha * sum_seg sum_k A_seg[m + shift, k] * B_seg[n, k]  (+ bias, + per-batch bias, + residual)
//
// One CTA per 128 x BN output tile.  Warp roles (192 threads for tiles <= 128 wide, 320 above):
//   warp 0   TMA producer: cp.async.bulk.tensor.4d into a ring of SWIZZLE_128B stages
//   warp 1   MMA issuer:   one thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32 in TMEM)
//   warp 2.. epilogue:     tcgen05.ld 32x32b -> registers -> fused epilogue -> global; one or two warps per TMEM lane quarter
// Operands may be K-major or MN-major (UMMA descriptors handle the transposed case), so the same kernel
// serves forward (K-major x K-major), data-gradient (K-major x MN-major) and weight-gradient
// (MN-major x MN-major, contraction over rows and batch) GEMMs, k=3 convolutions as three shifted
// segments (TMA out-of-bounds zero fill = conv padding), and the batched attention contractions.
#include <stdlib.h>

#include <mutex>

#include "tc_common.cuh"

namespace tc {
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}
}  // namespace tc

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct alignas(64) KParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  CUtensorMap tmO;      // bf16 output as (N, M, z2, z3), box 32 x 32, SWIZZLE_64B: the epilogue stores through TMA
  pt_segment_t seg[8];
  int nseg;
  int a_kmajor, b_kmajor;
  int a_batched[2], b_batched[2];
  int M, N, nz2, nz3;
  int mt, nt, tiles, work, streamk;   // mt counts groups of MC row tiles when the kernel runs as clusters of MC CTAs
  int gm;                             // raster: row-tile groups of gm walk all column tiles before the next group starts
  int total_iters;
  void* out;
  int out_dtype;
  long long osm, osz2, osz3;
  float alpha;
  const float* bias;
  const float* bias_z2;
  const bf16* res;
  long long rsm, rsz2, rsz3;
  long long bz2_stride;
  long long osn;
  int out_t;            // bf16 output / residual indexed [z][n][m] (M contiguous), bias per row: see pt_gemm_t.out_transposed
  int vec_red;          // fp32 atomic output: rows are contiguous and 16-byte aligned -> red.global.add.v4.f32
};

// PAIR: the kernel runs as CTA pairs (tcgen05 cta_group::2): one 256 x BN tile per pair, each CTA holds its own 128 rows of A and
// HALF of the B tile (the pair's MMA reads both halves), i.e. 32 KB instead of 48 KB of operands per SM and k-iteration at BN = 256.
template <int BN, bool PAIR = false>
struct Cfg {
  static constexpr int BN_LOAD = PAIR ? BN / 2 : BN;     // B rows (N) this CTA loads per stage
  static constexpr int BN_S = (BN_LOAD + 63) / 64 * 64;  // smem rows reserved for B (MN-major needs whole 64-blocks)
  static constexpr int B_STAGE_BYTES = BN_S * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int CTAS_PER_SM = (BN <= 128 && !PAIR) ? 2 : 1;  // narrow tiles: two persistent CTAs per SM hide each other's issue latency
  // Epilogue warps.  One warp per TMEM lane quarter is a single dependent instruction stream per SM sub-partition: ~1700
  // instructions per 128 x 256 tile at an IPC of ~0.13 = 13k clocks, more than the mainloop of any K < 1000 (ncu on
  // 24064 x 2560 x 320: same 87 us with K = 64 as with K = 320).  Wide tiles (one CTA per SM) therefore get TWO warps per
  // quarter, each taking every other 32-column pass; narrow tiles already have two CTAs = eight epilogue warps per SM.
  static constexpr int EPI_WARPS = CTAS_PER_SM == 1 ? 8 : 4;
  static constexpr int EPI_SPLIT = EPI_WARPS / 4;        // warps sharing one lane quarter
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int EPI_COLS = 32;                     // columns per epilogue pass
#ifndef PT_EPI_BUFS
#define PT_EPI_BUFS 1
#endif
  // Store tiles per warp.  Two (pass k+1 staged while TMA still reads the tile of pass k) were measured against one in a same-box
  // A/B of the whole step: 45.39 / 45.20 vs 45.27 / 45.46 ms -- no difference, so the default keeps the shared memory for the ring.
  static constexpr int EPI_BUFS = CTAS_PER_SM == 1 ? PT_EPI_BUFS : 1;
  static constexpr int EPI_BF16_BYTES = EPI_WARPS * EPI_BUFS * 32 * EPI_COLS * 2;   // dense 32 x 32 bf16 store tiles
  static constexpr int EPI_F32_BYTES = EPI_WARPS * 32 * 17 * 4;   // fp32 transpose tile per warp: 32 x 16 (+1), or 32 x 16 swizzled
  static constexpr int EPI_BYTES = EPI_BF16_BYTES > EPI_F32_BYTES ? EPI_BF16_BYTES : EPI_F32_BYTES;
  static constexpr int F32_COLS = 16;   // columns per transpose pass of the atomic epilogue
  static constexpr int BAR_BYTES = 512;     // keeps the epilogue staging 512-byte aligned (period of the 64-byte swizzle)
  static constexpr int BIAS_BYTES = 1024;
  static constexpr int AUX_BYTES = 1024 /*align slack*/ + BAR_BYTES + BIAS_BYTES + EPI_BYTES;
  static constexpr int STAGES_FIT = (227 * 1024 / CTAS_PER_SM - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES;
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // two accumulators
};

// Work distribution of the persistent kernel.  Data-parallel: CTA c takes output tiles c, c+G, ... whole.
// Stream-K (fp32 atomic-accumulate outputs only): the (tile, k-iteration) space is cut into G equal contiguous ranges,
// so every SM gets the same number of MMA iterations whatever the tile count; a range that crosses a tile boundary
// yields one segment per tile, each flushed with atomic adds.
struct Sched {
  int total_iters, streamk, cursor, end, stride;
  // `mc` CTAs of a cluster walk the same schedule in lockstep (they share B tiles by TMA multicast)
  __device__ __forceinline__ Sched(const KParams& p, int mc) {
    total_iters = p.total_iters;
    streamk = p.streamk;
    const int nclu = (int)gridDim.x / mc, clu = (int)blockIdx.x / mc;
    if (streamk) {
      const int per = (p.work + nclu - 1) / nclu;
      cursor = min(p.work, clu * per);
      end = min(p.work, cursor + per);
      stride = 0;
    } else {
      cursor = clu;
      end = p.tiles;
      stride = nclu;
    }
  }
  __device__ __forceinline__ bool next(int& tile, int& it_begin, int& it_end) {
    if (cursor >= end) return false;
    if (streamk) {
      tile = cursor / total_iters;
      it_begin = cursor - tile * total_iters;
      it_end = min(total_iters, it_begin + (end - cursor));
      cursor += it_end - it_begin;
    } else {
      tile = cursor;
      it_begin = 0;
      it_end = total_iters;
      cursor += stride;
    }
    return true;
  }
};

// MC = 1: independent CTAs.  MC = 2: clusters of two CTAs work on vertically adjacent 128-row tiles of the same column block;
// each loads its own A tile and HALF of the B tile, multicast into both CTAs' shared memory, which halves the B traffic from L2
// (operand delivery, ~64 B/clk/SM, is what limits the 256-wide tiles).  A stage is reusable when BOTH CTAs' MMAs have retired.
// PAIR (MC = 2): the two CTAs of a cluster are a tcgen05 CTA pair working on ONE 256-row tile.  Rank 0 (the leader) issues
// tcgen05.mma.cta_group::2; every CTA's TMA loads signal the LEADER's full barrier; stage release and accumulator-ready are
// multicast commits to both CTAs; both epilogues hand the accumulator back on the leader's tmem_empty barrier.
template <int BN, int MC, bool PAIR>
__global__ void __launch_bounds__(Cfg<BN, PAIR>::THREADS, Cfg<BN, PAIR>::CTAS_PER_SM) gemm_kernel(const __grid_constant__ KParams p) {
  using C = Cfg<BN, PAIR>;
  static_assert(!PAIR || MC == 2, "a CTA pair is a cluster of two");
  const int crank = MC > 1 ? (int)cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2] ; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
  float* sbias = reinterpret_cast<float*>(sgen + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES);
  uint8_t* sepi = sgen + C::STAGES * C::STAGE_BYTES + C::BAR_BYTES + C::BIAS_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA[p.seg[0].a_idx])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB[p.seg[0].b_idx])) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), PAIR ? 1 : MC);      // one tcgen05.commit arrival per issuing CTA (PAIR: the leader's multicast commit)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), PAIR ? 2 * C::EPI_WARPS : 32 * C::EPI_WARPS);   // PAIR: one elected lane per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MC > 1) cluster_sync_all();      // the peer's barriers are initialised before anything is multicast to them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(sgen + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));

  // tile index -> (m0, n0, z2, z3).  Within one (z2, z3) the tiles are walked in groups of `gm` row tiles: m fastest inside the
  // group (concurrently running CTAs share one B = weight tile through L2), then the next column tile of the SAME rows, and only
  // then the next group.  One wave of CTAs therefore covers (gm rows) x (all columns): every A row panel is fetched from DRAM once
  // and reused from L2 by the CTAs of its columns.  With m fastest over ALL rows (the first version) an A operand larger than L2
  // (24064 x 2560, 6016 x 10240: 123 MB) came from DRAM once per column tile: ncu on 6016 x 1280 x 10240 read 276 MB for 165 MB
  // of operands.
  auto decode = [&](int tile, int& m0, int& n0, int& z2, int& z3) {
    const int per_z = p.mt * p.nt;
    int r = tile / per_z;
    const int t = tile - r * per_z;
    const int span = p.gm * p.nt;                 // tiles of one full group
    const int grp = t / span;
    const int first = grp * p.gm;
    const int gsz = min(p.gm, p.mt - first);      // the last group may be shorter
    const int q = t - grp * span;
    const int mt = first + q % gsz;
    const int nt = q / gsz;
    m0 = (mt * MC + crank) * BM;
    n0 = nt * BN;
    z2 = r % p.nz2;
    z3 = r / p.nz2;
  };

  Sched sched(p, MC);
  int tile, it_begin, it_end;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform control flow, one elected lane issues)
    const bool leader = elect_one();
    const uint32_t tx_own = A_STAGE_BYTES + (p.b_kmajor ? C::BN_LOAD * BK * 2 : C::B_STAGE_BYTES);
    const uint32_t tx_bytes = PAIR ? 2 * tx_own : tx_own;        // PAIR: the leader's barrier counts the bytes of both CTAs
    const uint32_t full0_leader = PAIR ? mapa_u32(full_bar(0), 0) : 0;   // shared::cluster address of the leader's full[0]
    int s = 0;
    uint32_t ph = 1;
    while (sched.next(tile, it_begin, it_end)) {
      int m0, n0, z2, z3;
      decode(tile, m0, n0, z2, z3);
      // locate (seg, rep, kb) of it_begin
      int seg = 0, rep = 0, kb = 0, skip = it_begin;
      while (true) {
        const int nkb = (p.seg[seg].nk + BK - 1) / BK;
        const int cnt = nkb * p.seg[seg].nrep;
        if (skip < cnt) {
          rep = skip / nkb;
          kb = skip % nkb;
          break;
        }
        skip -= cnt;
        ++seg;
      }
      for (int it = it_begin; it < it_end; ++it) {
        mbar_wait(empty_bar(s), ph);
        const pt_segment_t& sg = p.seg[seg];
        if (PAIR && leader) {
          if (crank == 0) mbar_expect_tx(full_bar(s), tx_bytes);
          const uint32_t sa = smem_base + s * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint32_t fb = full0_leader + 8u * s;
          const CUtensorMap* ta = &p.tmA[sg.a_idx];
          const CUtensorMap* tb = &p.tmB[sg.b_idx];
          const int bz2 = sg.rep_is_batch ? (sg.rep_c2_0 + rep) : z2;
          const int a2 = p.a_batched[sg.a_idx] ? bz2 : 0, a3 = p.a_batched[sg.a_idx] ? z3 : 0;
          const int b2 = p.b_batched[sg.b_idx] ? bz2 : 0, b3 = p.b_batched[sg.b_idx] ? z3 : 0;
          const int ka = sg.a_k0 + kb * BK, kbb = sg.b_k0 + z2 * sg.b_k0_z2 + kb * BK;
          const int nb0 = n0 + sg.b_mn_shift + crank * C::BN_LOAD;     // this CTA's half of the B tile
          if (p.a_kmajor) {
            tma_load_4d_pair(sa, ta, fb, ka, m0 + sg.a_mn_shift, a2, a3);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d_pair(sa + j * (BK * 128), ta, fb, m0 + sg.a_mn_shift + j * 64, ka, a2, a3);
          }
          if (p.b_kmajor) {
            tma_load_4d_pair(sb, tb, fb, kbb, nb0, b2, b3);
          } else {
#pragma unroll
            for (int j = 0; j < C::BN_S / 64; ++j) tma_load_4d_pair(sb + j * (BK * 128), tb, fb, nb0 + j * 64, kbb, b2, b3);
          }
        } else if (leader) {
          mbar_expect_tx(full_bar(s), tx_bytes);
          const uint32_t sa = smem_base + s * C::STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE_BYTES;
          const CUtensorMap* ta = &p.tmA[sg.a_idx];
          const CUtensorMap* tb = &p.tmB[sg.b_idx];
          const int bz2 = sg.rep_is_batch ? (sg.rep_c2_0 + rep) : z2;
          const int a2 = p.a_batched[sg.a_idx] ? bz2 : 0, a3 = p.a_batched[sg.a_idx] ? z3 : 0;
          const int b2 = p.b_batched[sg.b_idx] ? bz2 : 0, b3 = p.b_batched[sg.b_idx] ? z3 : 0;
          const int ka = sg.a_k0 + kb * BK, kbb = sg.b_k0 + z2 * sg.b_k0_z2 + kb * BK;
          if (p.a_kmajor) {
            tma_load_4d(sa, ta, full_bar(s), ka, m0 + sg.a_mn_shift, a2, a3);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * (BK * 128), ta, full_bar(s), m0 + sg.a_mn_shift + j * 64, ka, a2, a3);
          }
          if (MC == 1) {
            if (p.b_kmajor) {
              tma_load_4d(sb, tb, full_bar(s), kbb, n0 + sg.b_mn_shift, b2, b3);
            } else {
#pragma unroll
              for (int j = 0; j < C::BN_S / 64; ++j) tma_load_4d(sb + j * (BK * 128), tb, full_bar(s), n0 + sg.b_mn_shift + j * 64, kbb, b2, b3);
            }
          } else {
            // this CTA fetches its half of the B tile and multicasts it to both CTAs of the pair (tensor map box = BN / 2 rows)
            constexpr int HB = BN / MC;
            if (p.b_kmajor) {
              tma_load_4d_mc(sb + crank * (HB * 128), tb, full_bar(s), kbb, n0 + sg.b_mn_shift + crank * HB, b2, b3, (uint16_t)((1 << MC) - 1));
            } else {
#pragma unroll
              for (int j = 0; j < HB / 64; ++j) {
                const int jj = crank * (HB / 64) + j;
                tma_load_4d_mc(sb + jj * (BK * 128), tb, full_bar(s), n0 + sg.b_mn_shift + jj * 64, kbb, b2, b3, (uint16_t)((1 << MC) - 1));
              }
            }
          }
        }
        // advance (seg, rep, kb) and the ring position
        const int nkb = (sg.nk + BK - 1) / BK;
        if (++kb == nkb) {
          kb = 0;
          if (++rep == sg.nrep) {
            rep = 0;
            ++seg;
          }
        }
        if (++s == C::STAGES) s = 0, ph ^= 1;
      }
    }
  } else if (warp == 1 && (!PAIR || crank == 0)) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control flow, one elected lane issues;
    // PAIR: only in the leader CTA, for both SMs)
    const bool leader = elect_one();
    // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
    const uint32_t idesc = idesc_f16(p.a_kmajor ? 0 : 1, p.b_kmajor ? 0 : 1, BN, PAIR ? 2 * BM : BM);
    const uint32_t a_lbo = p.a_kmajor ? 0u : (uint32_t)(BK * 128), b_lbo = p.b_kmajor ? 0u : (uint32_t)(BK * 128);
    const uint32_t a_kstep = p.a_kmajor ? 2u : 128u, b_kstep = p.b_kmajor ? 2u : 128u;  // (bytes >> 4) per UMMA_K=16
    // UMMA descriptors are linear in the shared-memory address: bases once, (bytes >> 4) added per use
    const uint64_t adesc0 = umma_desc(smem_base, a_lbo, 1024);
    const uint64_t bdesc0 = umma_desc(smem_base + A_STAGE_BYTES, b_lbo, 1024);
    int s = 0;
    uint32_t ph = 0;
    int segi = 0;
    while (sched.next(tile, it_begin, it_end)) {
      const int acc = segi & 1;
      mbar_wait(tmem_empty_bar(acc), ((segi >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t dcol = tmem_base + (uint32_t)(acc * BN);
      for (int it = it_begin; it < it_end; ++it) {
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          const uint64_t ad = adesc0 + (uint64_t)(s * (C::STAGE_BYTES >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(s * (C::STAGE_BYTES >> 4));
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_pair(dcol, ad + (uint64_t)(k * a_kstep), bd + (uint64_t)(k * b_kstep), idesc, (it > it_begin || k > 0) ? 1u : 0u);
            umma_commit_pair(empty_bar(s), (uint16_t)3);                          // both CTAs' stages are free when these MMAs retire
            if (it == it_end - 1) umma_commit_pair(tmem_full_bar(acc), (uint16_t)3);  // both epilogues may read their 128 rows
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16(dcol, ad + (uint64_t)(k * a_kstep), bd + (uint64_t)(k * b_kstep), idesc, (it > it_begin || k > 0) ? 1u : 0u);
            if (MC == 1) umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
            else umma_commit_mc(empty_bar(s), (uint16_t)((1 << MC) - 1));   // ... in both CTAs: the peer multicasts into this stage too
            if (it == it_end - 1) umma_commit(tmem_full_bar(acc));
          }
        }
        __syncwarp();
        if (++s == C::STAGES) s = 0, ph ^= 1;
      }
      ++segi;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (warps 2..), overlapped with the next segment's mainloop
    // hand an accumulator buffer back to the MMA issuer (PAIR: the leader's barrier, one elected lane per warp of both CTAs)
    const uint32_t tmem_empty_leader0 = PAIR ? mapa_u32(tmem_empty_bar(0), 0) : 0;
    auto release_acc = [&](int a) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tmem_empty_leader0 + 8u * a);
      } else {
        mbar_arrive(tmem_empty_bar(a));
      }
    };
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int ew = warp - 2;             // epilogue warp index: private staging tile
    const int half = ew >> 2;            // which of the EPI_SPLIT warps of this quarter: takes passes half, half + EPI_SPLIT, ...
    constexpr int ES = C::EPI_SPLIT;
    constexpr int ET = 32 * C::EPI_WARPS;
    const int et = threadIdx.x - 64;
    int segi = 0;
    int pc = 0;                          // bf16 path: store passes done by this warp (selects the staging tile)
    int staged_n0 = -1, staged_z2 = -1;
    while (sched.next(tile, it_begin, it_end)) {
      int m0, n0, z2, z3;
      decode(tile, m0, n0, z2, z3);
      const int acc = segi & 1;
      // stage the per-column additive term (bias + per-batch time shift) in shared memory -- only when the column block (or the
      // batch element of a time shift) differs from what is already there: tiles are walked m-fastest, so this is rare
      if (!p.out_t && (n0 != staged_n0 || (p.bias_z2 != nullptr && z2 != staged_z2))) {
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");   // everyone is done with the previous contents
        const float* bz = p.bias_z2 ? p.bias_z2 + (long long)z2 * p.bz2_stride : nullptr;
        for (int j = et; j < BN; j += ET) {
          const int n = n0 + j;
          float bsum = 0.f;
          if (n < p.N) {
            if (p.bias) bsum += __ldg(p.bias + n);
            if (bz) bsum += __ldg(bz + n);
          }
          sbias[j] = bsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
        staged_n0 = n0;
        staged_z2 = z2;
      }
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const uint32_t full_parity = (segi >> 1) & 1;
      if (p.out_dtype == PT_OUT_BF16) {
        // ---- bf16 output (+ optional bf16 residual): every global access is a full 64/128-byte row segment.
        // Accumulator rows live one per thread; a warp-private padded smem tile transposes between "thread = row"
        // and "8 (or 4) lanes = one contiguous row segment".
        constexpr int CH = C::EPI_COLS;            // columns per pass (32: one 64-byte row segment)
        constexpr int NPASS = BN / CH;
        constexpr int NPH = (NPASS + ES - 1) / ES; // passes per warp
        constexpr int VPR = CH / 8;                // 16-byte vectors per row segment (4)
        constexpr int RPI = 32 / VPR;              // rows covered by one warp-wide residual load
        constexpr int NIT = 32 / RPI;              // residual loads per pass
        // Store tile of this warp: 32 rows x 64 bytes, dense, in the SWIZZLE_64B pattern of the output tensor map (16-byte chunk
        // index ^= address bits 7-8), which also makes the row-per-thread writes below bank-conflict free.  One elected lane hands
        // the tile to TMA: no LDS / STG / 64-bit address arithmetic / tail predicates in the epilogue (tails are clipped by TMA).
        // Two tiles per warp, used alternately (`pc` = passes this warp has stored so far): waiting for TMA to finish READING the
        // single tile before every pass serialised the epilogue with the store latency (4 passes x ~1 us per 128 x 256 tile, more
        // than the whole mainloop of a K <= 640 GEMM).
        uint8_t* stg_base = sepi + ew * (C::EPI_BUFS * 32 * CH * 2);
        uint8_t* stg = stg_base + (pc & (C::EPI_BUFS - 1)) * (32 * CH * 2);
        auto cell = [&](int row, int chunk) { return stg + row * (CH * 2) + ((chunk ^ ((row >> 1) & 3)) << 4); };
        const int mw = m0 + q * 32;                // first row of this warp
        const int crow = lane / VPR, cvec = lane % VPR;
        const long long zoff_r = (long long)z2 * p.rsz2 + (long long)z3 * p.rsz3;
        const bool has_res = p.res != nullptr;
        const int last_c = min(NPASS, (p.N - n0 + CH - 1) / CH) - 1;   // last pass with columns inside N
        const int last_k = last_c >= half ? (last_c - half) / ES : -1;  // this warp's last pass (pass index = half + k * ES)
        uint32_t v[2][CH];
        bf16x8 rr[2][NIT];
        auto load_res = [&](int pass, bf16x8* dst) {
          if (p.out_t) {      // staging tile = [32 columns n][32 rows m]: the residual is read along its contiguous m
            const int mb = mw + cvec * 8;
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
              const int n = n0 + pass * CH + crow + i * RPI;
              if (n < p.N && mb < p.M) dst[i] = ld16(p.res + zoff_r + (long long)n * p.rsm + mb);
            }
            return;
          }
          const int nb = n0 + pass * CH + cvec * 8;
#pragma unroll
          for (int i = 0; i < NIT; ++i) {
            const int m = mw + crow + i * RPI;
            if (m < p.M && nb < p.N) dst[i] = ld16(p.res + zoff_r + (long long)m * p.rsm + nb);
          }
        };
        float brow = 0.f;     // transposed output: the additive term belongs to this thread's row
        if (p.out_t && mw + lane < p.M) {
          if (p.bias) brow += __ldg(p.bias + mw + lane);
          if (p.bias_z2) brow += __ldg(p.bias_z2 + (long long)z2 * p.bz2_stride + mw + lane);
        }
        if (has_res && last_k >= 0) load_res(half, rr[0]);
        mbar_wait(tmem_full_bar(acc), full_parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (last_k < 0) {   // no column of this warp's passes is inside N: nothing to read, hand the buffer back
          release_acc(acc);
        } else {
          tmem_ld32(tbase + (uint32_t)(half * CH), v[0]);
        }
#pragma unroll
        for (int k = 0; k < NPH; ++k) {
          const int c = half + k * ES;             // pass index (warp-uniform)
          const int nb = n0 + c * CH;
          if (k <= last_k) {  // warp-uniform
            __syncwarp();
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (k < last_k) {
              tmem_ld32(tbase + (uint32_t)((c + ES) * CH), v[(k + 1) & 1]);
              if (has_res) load_res(c + ES, rr[(k + 1) & 1]);
            } else {
              // the accumulator is in registers: hand the TMEM buffer back to the MMA warp
              release_acc(acc);
            }
            stg = stg_base + (pc & (C::EPI_BUFS - 1)) * (32 * CH * 2);
            if (lane == 0) {   // the store that last used THIS tile has read it
              if (C::EPI_BUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            if (has_res) {   // residual rows arrive coalesced (8 rows x 64 bytes per access) and are re-read row-per-thread
#pragma unroll
              for (int i = 0; i < NIT; ++i) st16(cell(crow + i * RPI, cvec), rr[k & 1][i]);
              __syncwarp();
            }
            if (p.out_t) {
              // the tile leaves as [32 columns n][32 rows m] (m contiguous in memory): this thread owns column `lane` of every staging
              // row -- 32 two-byte accesses instead of 4 sixteen-byte ones (consecutive lanes hit consecutive bf16: no bank conflicts)
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                bf16* cellp = reinterpret_cast<bf16*>(cell(j, lane >> 3)) + (lane & 7);
                float f = fmaf(__uint_as_float(v[k & 1][j]), p.alpha, brow);
                if (has_res) f += __bfloat162float(*cellp);
                *cellp = __float2bfloat16_rn(f);
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&p.tmO, smem_u32(stg), mw, nb, z2, z3);
                tma_store_commit();
              }
              ++pc;
              continue;
            }
            const float* sb = sbias + c * CH;
#pragma unroll
            for (int g = 0; g < VPR; ++g) {
              const float4 b0 = *reinterpret_cast<const float4*>(sb + g * 8), b1 = *reinterpret_cast<const float4*>(sb + g * 8 + 4);
              const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(__uint_as_float(v[k & 1][g * 8 + j]), p.alpha, bs[j]);
              uint8_t* mine = cell(lane, g);       // this thread's row, vector g
              if (has_res) {
                const bf16x8 t = ld16(mine);
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  const float2 u = bf2_to_f2(t.w[h]);
                  f[2 * h] += u.x;
                  f[2 * h + 1] += u.y;
                }
              }
              bf16x8 t;
#pragma unroll
              for (int h = 0; h < 4; ++h) t.w[h] = f2_to_bf2(f[2 * h], f[2 * h + 1]);
              st16(mine, t);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&p.tmO, smem_u32(stg), nb, mw, z2, z3);
              tma_store_commit();
            }
            ++pc;
          }
        }
      } else {
        // ---- fp32 store / fp32 atomic accumulate (tiny MLP outputs, weight gradients): one row per thread, 16 columns per pass
        const int m = m0 + q * 32 + lane;
        const bool m_ok = m < p.M;
        const long long out_off = (long long)z2 * p.osz2 + (long long)z3 * p.osz3 + (long long)m * p.osm;
        constexpr int FC = C::F32_COLS;            // 16
        constexpr int NCH = BN / FC;
        constexpr int NPH = (NCH + ES - 1) / ES;   // chunks per warp
        const int last_c = min(NCH, (p.N - n0 + FC - 1) / FC) - 1;
        const int last_k = last_c >= half ? (last_c - half) / ES : -1;   // this warp's last chunk (chunk index = half + k * ES)
        uint32_t v[2][FC];
        mbar_wait(tmem_full_bar(acc), full_parity);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (last_k < 0) {
          release_acc(acc);
        } else {
          tmem_ld16(tbase + (uint32_t)(half * FC), v[0]);
        }
        float* tile = reinterpret_cast<float*>(sepi) + ew * (32 * (FC + 1));
        const int mrow0 = m0 + q * 32;
        const long long zoff = (long long)z2 * p.osz2 + (long long)z3 * p.osz3;
#pragma unroll
        for (int k = 0; k < NPH; ++k) {
          const int c = half + k * ES;
          const int nb = n0 + c * FC;
          if (k <= last_k) {  // warp-uniform
            __syncwarp();
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (k < last_k) {
              tmem_ld16(tbase + (uint32_t)((c + ES) * FC), v[(k + 1) & 1]);
            } else {
              release_acc(acc);
            }
            float f[FC];
#pragma unroll
            for (int j = 0; j < FC; ++j) f[j] = fmaf(__uint_as_float(v[k & 1][j]), p.alpha, sbias[c * FC + j]);
            if (p.out_dtype == PT_OUT_F32) {
              if (m_ok) {
                float* o = reinterpret_cast<float*>(p.out) + out_off + nb;
#pragma unroll
                for (int g = 0; g < FC / 4; ++g) {
                  if (nb + g * 4 + 4 <= p.N) {
                    *reinterpret_cast<float4*>(o + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      if (nb + g * 4 + j < p.N) o[g * 4 + j] = f[g * 4 + j];
                  }
                }
              }
            } else if (p.vec_red) {
              // PT_OUT_F32_ATOMIC_ADD, contiguous 16-byte aligned rows: transpose through a warp-private 32 x 16 tile (16-byte chunks
              // XOR-swizzled by the row pair: conflict-free both ways) and flush with red.global.add.v4.f32 -- four floats per lane
              // request, four lanes per 64-byte row segment, 8 rows per warp instruction (4 instructions per pass instead of 16)
              float* vt = reinterpret_cast<float*>(sepi) + ew * (32 * FC);
#pragma unroll
              for (int g = 0; g < FC / 4; ++g)
                *reinterpret_cast<float4*>(vt + lane * FC + ((g ^ ((lane >> 1) & 3)) << 2)) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
              __syncwarp();
              const int c4 = lane & 3, rsub = lane >> 2;
              const int n = nb + c4 * 4;
              if (n < p.N) {
                float* o = reinterpret_cast<float*>(p.out) + zoff + n;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                  const int r = it * 8 + rsub;
                  if (mrow0 + r < p.M) {
                    const float4 t = *reinterpret_cast<const float4*>(vt + r * FC + ((c4 ^ ((r >> 1) & 3)) << 2));
                    float* dst = o + (long long)(mrow0 + r) * p.osm;
                    if (n + 4 <= p.N) {
                      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
                    } else {
                      const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                      for (int j = 0; j < 4; ++j)
                        if (n + j < p.N) atomicAdd(dst + j, tv[j]);
                    }
                  }
                }
              }
              __syncwarp();
            } else {
              // PT_OUT_F32_ATOMIC_ADD, strided output: transpose through a warp-private tile so that one warp-wide RED covers
              // consecutive columns of two rows (8 floats per 32-byte L2 sector instead of 1)
#pragma unroll
              for (int j = 0; j < FC; ++j) tile[lane * (FC + 1) + j] = f[j];
              __syncwarp();
              const int col = lane % FC, rsub = lane / FC;
              const int n = nb + col;
              if (n < p.N) {
                float* o = reinterpret_cast<float*>(p.out) + zoff + (long long)n * p.osn;
#pragma unroll 8
                for (int r = rsub; r < 32; r += 32 / FC) {
                  if (mrow0 + r < p.M) atomicAdd(o + (long long)(mrow0 + r) * p.osm, tile[r * (FC + 1) + col]);
                }
              }
            }
          }
        }
      }
      ++segi;
    }
    if (lane == 0) tma_store_wait_all();   // this warp's output tiles have left shared memory and are written
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MC > 1) cluster_sync_all();      // do not leave while the peer can still multicast into / arrive on this CTA's shared memory
  if (warp == 2) {
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host
int encode_operand(CUtensorMap* tm, const pt_operand_t& op, int box_rows_kmajor, const char* name) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    pt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return PT_ECUDA;
  }
  PT_REQUIRE(op.ptr != nullptr, "gemm operand %s: null pointer", name);
  PT_REQUIRE((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, "gemm operand %s: base not 16-byte aligned", name);
  PT_REQUIRE(op.stride[0] == 1, "gemm operand %s: dim[0] must be contiguous", name);
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4] = {64, 1, 1, 1}, estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    PT_REQUIRE(op.dim[i] >= 1 && op.dim[i] < (1ll << 31), "gemm operand %s: dim[%d]=%lld out of range", name, i, (long long)op.dim[i]);
    dims[i] = (cuuint64_t)op.dim[i];
  }
  for (int i = 1; i < 4; ++i) {
    long long st = op.stride[i];
    if (op.dim[i] == 1 && (st <= 0 || (st % 8) != 0)) st = 8;  // irrelevant axis, keep the encoder happy
    PT_REQUIRE(st > 0 && st % 8 == 0, "gemm operand %s: stride[%d]=%lld must be a positive multiple of 8 elements", name, i, st);
    strides[i - 1] = (cuuint64_t)st * 2;
  }
  box[1] = op.kmajor ? (cuuint32_t)box_rows_kmajor : 64u;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pt_set_error("cuTensorMapEncodeTiled(%s) failed: CUresult %d (dims %lld,%lld,%lld,%lld strides %lld,%lld,%lld box %u,%u)", name, (int)r,
                 (long long)dims[0], (long long)dims[1], (long long)dims[2], (long long)dims[3], (long long)strides[0],
                 (long long)strides[1], (long long)strides[2], box[0], box[1]);
    return PT_ECUDA;
  }
  return PT_OK;
}

// bf16 output as a rank-4 tensor (N, M, z2, z3) for the epilogue's TMA stores: box 32 x 32, SWIZZLE_64B
int encode_output(CUtensorMap* tm, const pt_gemm_t* g) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    pt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return PT_ECUDA;
  }
  PT_REQUIRE((reinterpret_cast<uintptr_t>(g->out) & 15) == 0, "pt_gemm: bf16 output base not 16-byte aligned");
  // transposed output: the contiguous axis is M and out_stride_m is the stride of the N index
  const long long ext[4] = {g->out_transposed ? g->M : g->N, g->out_transposed ? g->N : g->M, g->nz2, g->nz3};
  const long long str[4] = {1, g->out_stride_m, g->out_stride_z2, g->out_stride_z3};
  cuuint64_t dims[4], strides[3];
  for (int i = 0; i < 4; ++i) dims[i] = (cuuint64_t)ext[i];
  for (int i = 1; i < 4; ++i) {
    long long st = str[i];
    if (ext[i] == 1 && (st <= 0 || (st % 8) != 0)) st = 8;  // irrelevant axis, keep the encoder happy
    PT_REQUIRE(st > 0 && st % 8 == 0, "pt_gemm: bf16 output stride[%d]=%lld must be a positive multiple of 8 elements", i, st);
    strides[i - 1] = (cuuint64_t)st * 2;
  }
  cuuint32_t box[4] = {32, 32, 1, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, g->out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pt_set_error("cuTensorMapEncodeTiled(out) failed: CUresult %d (dims %lld,%lld,%lld,%lld strides %lld,%lld,%lld)", (int)r, ext[0], ext[1],
                 ext[2], ext[3], (long long)strides[0], (long long)strides[1], (long long)strides[2]);
    return PT_ECUDA;
  }
  return PT_OK;
}

template <int BN, int MC, bool PAIR = false>
int launch(const KParams& kp, dim3 grid, cudaStream_t st) {
  using CF = Cfg<BN, PAIR>;
  PT_ONCE_PER_DEVICE(cudaFuncSetAttribute(gemm_kernel<BN, MC, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
  if (MC == 1) {
    gemm_kernel<BN, MC, PAIR><<<grid, CF::THREADS, CF::SMEM_BYTES, st>>>(kp);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(CF::THREADS, 1, 1);
    cfg.dynamicSmemBytes = CF::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = MC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // the schedule strides by the number of clusters launched, so never launch more than can be resident at once (a GPC whose SM
    // count is not a multiple of the cluster size leaves SMs unused)
    static int max_clusters = 0;
    if (!max_clusters) {
      int n = 0;
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3((unsigned)(pt_num_sms_physical() / MC * MC), 1, 1);
      if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<BN, MC, PAIR>, &q) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = pt_num_sms_physical() / MC;
      }
      max_clusters = n;
      if (getenv("PT_GEMM_DEBUG")) fprintf(stderr, "[pt_gemm] BN=%d cluster=%d pair=%d: %d clusters resident at once\n", BN, MC, (int)PAIR, n);
    }
    if ((int)grid.x / MC > max_clusters) cfg.gridDim = dim3((unsigned)(max_clusters * MC), 1, 1);
    PT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, MC, PAIR>, kp));
  }
  PT_LAUNCH_CHECK();
  return PT_OK;
}

}  // namespace

static thread_local int g_last_tile = 0;
extern "C" int pt_gemm_last_tile(void) { return g_last_tile; }

extern "C" int pt_gemm(const pt_gemm_t* g, void* stream) {
  PT_REQUIRE(g != nullptr, "pt_gemm: null descriptor");
  PT_REQUIRE(g->nseg >= 1 && g->nseg <= 8, "pt_gemm: nseg=%d", g->nseg);
  PT_REQUIRE(g->M >= 1 && g->N >= 1, "pt_gemm: M=%d N=%d", g->M, g->N);
  PT_REQUIRE(g->nz2 >= 1 && g->nz3 >= 1, "pt_gemm: bad batch extents");
  PT_REQUIRE(g->out != nullptr, "pt_gemm: null output");
  if (g->out_dtype == PT_OUT_BF16) {
    PT_REQUIRE((g->out_transposed ? g->M : g->N) % 8 == 0 && g->out_stride_m % 8 == 0 && g->out_stride_z2 % 8 == 0 && g->out_stride_z3 % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(g->out) & 15) == 0,
               "pt_gemm: bf16 output needs N (M if transposed), strides multiples of 8 and a 16-byte aligned base (M=%d N=%d)", g->M, g->N);
  } else if (g->out_dtype == PT_OUT_F32) {
    PT_REQUIRE(g->out_stride_m % 4 == 0 && g->out_stride_z2 % 4 == 0 && g->out_stride_z3 % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(g->out) & 15) == 0,
               "pt_gemm: f32 output needs N, strides multiples of 4 (N=%d)", g->N);
  } else {
    PT_REQUIRE(g->out_dtype == PT_OUT_F32_ATOMIC_ADD, "pt_gemm: out_dtype=%d", g->out_dtype);
  }
  if (g->residual) {
    PT_REQUIRE((g->out_transposed ? g->M : g->N) % 8 == 0 && g->res_stride_m % 8 == 0 && g->res_stride_z2 % 8 == 0 && g->res_stride_z3 % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0,
               "pt_gemm: residual alignment");
  }

  KParams kp;
  memset(&kp, 0, sizeof(kp));
  bool a_used[2] = {false, false}, b_used[2] = {false, false};
  long long total = 0;
  for (int i = 0; i < g->nseg; ++i) {
    const pt_segment_t& s = g->seg[i];
    PT_REQUIRE(s.a_idx >= 0 && s.a_idx < 2 && s.b_idx >= 0 && s.b_idx < 2, "pt_gemm: segment %d operand index", i);
    PT_REQUIRE(s.nk >= 1 && s.nrep >= 1, "pt_gemm: segment %d nk=%d nrep=%d", i, s.nk, s.nrep);
    a_used[s.a_idx] = true;
    b_used[s.b_idx] = true;
    kp.seg[i] = s;
    total += (long long)((s.nk + BK - 1) / BK) * s.nrep;
  }
  PT_REQUIRE(total < (1ll << 30), "pt_gemm: too many k iterations");
  const int a0 = a_used[0] ? 0 : 1, b0 = b_used[0] ? 0 : 1;
  kp.a_kmajor = g->a[a0].kmajor;
  kp.b_kmajor = g->b[b0].kmajor;
  for (int i = 0; i < 2; ++i) {
    if (a_used[i]) PT_REQUIRE(g->a[i].kmajor == kp.a_kmajor, "pt_gemm: A operands must share one majorness");
    if (b_used[i]) PT_REQUIRE(g->b[i].kmajor == kp.b_kmajor, "pt_gemm: B operands must share one majorness");
  }

  // tile width: cost model fitted to tools/gemm_sweep.py on B200 (profiles/r01_gemm_sweep_v3.txt).  One k-iteration (BK = 64) of a
  // 128 x bn tile costs ~800 / 730 / 660 / 640 / 430 / 260 cycles for bn = 256 / 224 / 192 / 160 / 128 / 64 in this model (an
  // effective figure that folds the per-tile epilogue and fill in; the marginal mainloop cost is the tensor pipe's, see DESIGN.md 3.1;
  // what favours wide tiles is the fixed cost per tile).  Data-parallel launches are paced by the busiest SM (ceil(tiles / SMs)
  // tiles); stream-K launches are balanced but pay one atomic flush of a tile per segment (~0.65 cycles per flushed column-row),
  // and their 64-wide tiles (MN-major operands, two CTAs per SM sharing the flush path) run ~1.6x slower than that table.
  const bool streamk = g->out_dtype == PT_OUT_F32_ATOMIC_ADD;
  PT_REQUIRE(!streamk || (g->bias == nullptr && g->bias_z2 == nullptr && g->residual == nullptr),
             "pt_gemm: PT_OUT_F32_ATOMIC_ADD outputs are scheduled stream-K and take no bias/residual");
  int bn = g->block_n & ~1;                 // bit 0 of block_n: 1 = do not form clusters (calibration runs)
  bool table_no_cluster = false;
  const bool allow_cluster_arg = (g->block_n & 1) == 0;
  const long long mt = (g->M + BM - 1) / BM;
  const long long zz = (long long)g->nz2 * g->nz3;
  const int sms = pt_num_sms();
  if (g->block_n == 0) {
    // measured per-shape overrides of the cost model below (the shapes of the bench model's train step; PT_GEMM_NO_TABLE=1 ignores them)
    struct TileRow { int M, N, K, zz, nseg, ak, bk, od, res, bn; };
    static const TileRow table[] = {
#include "gemm_tile_table.inc"
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
    static const bool no_table = getenv("PT_GEMM_NO_TABLE") != nullptr;
    if (!no_table) {
      long long ktot = 0;
      for (int s = 0; s < g->nseg; ++s) ktot += (long long)g->seg[s].nk * g->seg[s].nrep;
      for (const TileRow& r : table)
        if (r.M == g->M && r.N == g->N && r.K == ktot && r.zz == zz && r.nseg == g->nseg && r.ak == kp.a_kmajor && r.bk == kp.b_kmajor &&
            r.od == g->out_dtype && r.res == (g->residual != nullptr)) {
          bn = r.bn & ~1;
          table_no_cluster = (r.bn & 1) != 0;
          break;
        }
    }
  }
  const bool allow_cluster = allow_cluster_arg && !table_no_cluster;
  if (bn == 0) {
    // 224 / 192 exist for wave quantisation: when a narrower tile keeps the number of waves, every wave gets shorter
    const int cand[6] = {256, 224, 192, 128, 64, 160};
    const double cyc[6] = {mt >= 2 && allow_cluster && !streamk ? 780.0 : 800.0, 730.0, 660.0, 430.0, streamk ? 416.0 : 260.0, 640.0};
    double best = 1e300;
    for (int i = 0; i < 6; ++i) {
      const int c = cand[i];
      const long long tiles_c = mt * ((g->N + c - 1) / c) * zz;
      double cost;
      if (streamk) {
        const double slots_c = (double)sms * (c <= 128 ? 2 : 1);
        cost = (double)tiles_c * (double)total * cyc[i] / sms + 0.65 * ((double)tiles_c + slots_c) * c;
      } else {
        cost = (double)((tiles_c + sms - 1) / sms) * ((double)total * cyc[i] + 500.0);
      }
      if (cost < best * 0.97) {   // ties go to the earlier (wider) candidate
        best = cost;
        bn = c;
      }
    }
  }
  PT_REQUIRE(bn == 64 || bn == 96 || bn == 128 || bn == 160 || bn == 192 || bn == 224 || bn == 256, "pt_gemm: block_n=%d", bn);
  PT_REQUIRE(bn != 96 || kp.b_kmajor, "pt_gemm: block_n = 96 needs a K-major B operand");
  PT_REQUIRE(!g->out_transposed || (g->out_dtype == PT_OUT_BF16 && g->M % 8 == 0), "pt_gemm: out_transposed needs a bf16 output and M %% 8 == 0 (M=%d)", g->M);

  // clusters of two CTAs (TMA multicast of the B tile) for the 256-wide tiles whenever there are two row tiles to pair
  // (measured: +1-3 % on the large data-parallel shapes, -4 % on stream-K ones, so only the former use it)
  // (a cluster of four sharing B in quarters was measured too: 24064 x 2560 x 320 67 -> 74 us, 6016 x 3840 x 1280 57 -> 64 us --
  // what limits the mainloop is each SM's own ingest rate, ~64 B/clk, which multicast does not change)
  // (pairing along the BATCH axis for launches with one row tile per sample -- k=3 convolutions over <= 128 rows -- so that two samples
  // share the weight tile was measured too: 56.3 vs 52.8 us at 32 x 94 x 1280 -> 1280, slower; profiles/r02_small_conv_probe.txt)
  const int mc = (bn == 256 && mt >= 2 && allow_cluster && !streamk) ? 2 : 1;
  // ... and those clusters run as tcgen05 CTA pairs (cta_group::2: one 256 x 256 tile per pair, each SM ingests 32 KB instead of
  // 48 KB per k-iteration) unless an odd, small number of row tiles would leave a quarter of a pair's work empty
  static const bool no_pair = getenv("PT_GEMM_NO_PAIR") != nullptr;
  // Measured per shape (tools/gemm_sweep.py, profiles/r02_gemm_sweep.txt): the pair wins 3-10 % from 16 k-iterations per tile up
  // (K >= 1024) and loses up to 20 % on short contractions, where the cross-SM commit / barrier round trip per tile is not amortised
  // (24064 x 2560 x 320: 62 vs 51 us).
  const bool pair = mc == 2 && !no_pair && (mt % 2 == 0 || mt >= 8) && total >= 16;
  for (int i = 0; i < 2; ++i) {
    if (a_used[i]) {
      int r = encode_operand(&kp.tmA[i], g->a[i], BM, i ? "A1" : "A0");
      if (r) return r;
      kp.a_batched[i] = g->a[i].batched;
    }
    if (b_used[i]) {
      int r = encode_operand(&kp.tmB[i], g->b[i], bn / mc, i ? "B1" : "B0");
      if (r) return r;
      kp.b_batched[i] = g->b[i].batched;
    }
  }
  if (g->out_dtype == PT_OUT_BF16) {
    if (int r = encode_output(&kp.tmO, g)) return r;
  }
  kp.nseg = g->nseg;
  kp.M = g->M;
  kp.N = g->N;
  kp.nz2 = g->nz2;
  kp.nz3 = g->nz3;
  kp.total_iters = (int)total;
  const long long nt = (g->N + bn - 1) / bn;
  const long long mtg = (mt + mc - 1) / mc;          // row-tile groups: one per cluster
  const long long tiles = mtg * nt * zz;
  PT_REQUIRE(tiles * total < (1ll << 31), "pt_gemm: work space too large");
  kp.mt = (int)mtg;
  kp.nt = (int)nt;
  kp.tiles = (int)tiles;
  kp.work = (int)(tiles * total);
  kp.streamk = streamk ? 1 : 0;
  {
    // rows per raster group: about one wave of resident CTAs (clusters) spread over all column tiles
    static const bool no_group = getenv("PT_GEMM_NO_GROUP") != nullptr;
    const long long resident = (long long)sms * (bn <= 128 ? 2 : 1) / mc;
    long long gm = (resident + nt / 2) / nt;
    if (gm < 1) gm = 1;
    if (gm > mtg || no_group || streamk) gm = mtg;
    kp.gm = (int)gm;
  }
  kp.out = g->out;
  kp.out_dtype = g->out_dtype;
  kp.osm = g->out_stride_m;
  kp.osz2 = g->out_stride_z2;
  kp.osz3 = g->out_stride_z3;
  kp.alpha = g->alpha;
  kp.bias = g->bias;
  kp.bias_z2 = g->bias_z2;
  kp.res = reinterpret_cast<const bf16*>(g->residual);
  kp.rsm = g->res_stride_m;
  kp.rsz2 = g->res_stride_z2;
  kp.rsz3 = g->res_stride_z3;
  kp.bz2_stride = g->bias_z2_stride ? g->bias_z2_stride : g->N;
  kp.osn = g->out_stride_n ? g->out_stride_n : 1;
  kp.out_t = g->out_transposed ? 1 : 0;
  PT_REQUIRE(kp.osn == 1 || g->out_dtype == PT_OUT_F32_ATOMIC_ADD, "pt_gemm: out_stride_n needs PT_OUT_F32_ATOMIC_ADD");
  kp.vec_red = (g->out_dtype == PT_OUT_F32_ATOMIC_ADD && kp.osn == 1 && g->out_stride_m % 4 == 0 && g->out_stride_z2 % 4 == 0 &&
                g->out_stride_z3 % 4 == 0 && (reinterpret_cast<uintptr_t>(g->out) & 15) == 0 && !getenv("PT_GEMM_SCALAR_RED"))
                   ? 1 : 0;

  // persistent grid: one CTA per SM for the wide tiles, two for the narrow ones (fewer when there is less work)
  const long long slots = (long long)sms * (bn <= 128 ? 2 : 1) / mc;     // clusters (or single CTAs) that can be resident
  long long G = streamk ? (kp.work / 4 > 0 ? kp.work / 4 : 1) : tiles;
  if (G > slots) G = slots;
  dim3 grid((unsigned)(G * mc), 1, 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  g_last_tile = bn | (bn == 256 && mc == 1 ? 1 : 0);
  switch (bn) {
    case 64: return launch<64, 1>(kp, grid, st);
    case 96: return launch<96, 1>(kp, grid, st);
    case 128: return launch<128, 1>(kp, grid, st);
    case 160: return launch<160, 1>(kp, grid, st);
    case 192: return launch<192, 1>(kp, grid, st);
    case 224: return launch<224, 1>(kp, grid, st);
    default: return mc == 2 ? (pair ? launch<256, 2, true>(kp, grid, st) : launch<256, 2>(kp, grid, st)) : launch<256, 1>(kp, grid, st);
  }
}
