// TEST INFRASTRUCTURE ONLY -- host-side index checker for prompt_tts_b200/csrc/seanet/seanet_core.h.
//
// The SEANet kernels are barrier-free "one thread = one register tile" kernels whose bodies and launch geometry are plain
// inline functions in seanet_core.h.  This file compiles that header with g++ and walks each kernel's grid with host loops
// (block by block, thread by thread, the exact (blockIdx, threadIdx, blockDim) the device would pass), on HOST pointers.
// tests/test_seanet_host.py compares the results with oracle/seanet_oracle.py, which checks the padding / stride / phase /
// packing arithmetic of the kernels in the CPU-only test tier.  It is not a CPU path of the product: nothing under
// prompt_tts_b200/ builds, loads or calls it, and the product raises when libpt_seanet.so or a GPU is missing.
//
// Build: g++ -O2 -fPIC -shared -o oracle/_build/libseanet_emul.so tests/seanet_emul.cpp   (oracle/Makefile)
#include <vector>

#include "../prompt_tts_b200/csrc/seanet/seanet_core.h"

template <typename F>
static void walk(const sn_grid& g, F body) {
  for (unsigned bz = 0; bz < g.z; ++bz)
    for (unsigned by = 0; by < g.y; ++by)
      for (unsigned bx = 0; bx < g.x; ++bx)
        for (int tx = 0; tx < SN_THREADS; ++tx) body((int)bx, (int)by, (int)bz, tx);
}

extern "C" {

void emu_sn_conv1d(const pt_sn_conv_t* p) {
  walk(sn_conv1d_grid(*p), [&](int bx, int by, int bz, int tx) { sn_conv1d_thread(*p, bx, by, bz, tx, SN_THREADS); });
}
void emu_sn_conv_transpose1d(const pt_sn_conv_t* p) {
  walk(sn_convtr_grid(*p), [&](int bx, int by, int bz, int tx) { sn_convtr_thread(*p, bx, by, bz, tx, SN_THREADS); });
}
void emu_sn_weight_norm_fold(const float* v, const float* g, float* w, int rows, int cols) {
  walk(sn_linear_grid_1d(rows), [&](int bx, int, int, int tx) { sn_weight_norm_thread(v, g, w, rows, cols, bx, tx, SN_THREADS); });
}
void emu_sn_lstm_pack(const float* w, float* wt4, int H) {
  walk(sn_linear_grid_1d(4LL * H * H), [&](int bx, int, int, int tx) { sn_lstm_pack_thread(w, wt4, H, bx, tx, SN_THREADS); });
}
void emu_sn_lstm_pack_bias(const float* b_ih, const float* b_hh, float* bias4, int H) {
  walk(sn_linear_grid_1d(4LL * H), [&](int bx, int, int, int tx) { sn_lstm_pack_bias_thread(b_ih, b_hh, bias4, H, bx, tx, SN_THREADS); });
}
void emu_sn_ncl_to_tbc(const float* x, float* out, int B, int Cn, int T) {
  walk(sn_linear_grid_1d((long long)B * Cn * T), [&](int bx, int, int, int tx) { sn_ncl_to_tbc_thread(x, out, B, Cn, T, bx, tx, SN_THREADS); });
}
void emu_sn_tbc_add_to_ncl(const float* hseq, const float* x, float* y, float* y_elu, int B, int Cn, int T) {
  walk(sn_linear_grid_1d((long long)B * Cn * T),
       [&](int bx, int, int, int tx) { sn_tbc_add_to_ncl_thread(hseq, x, y, y_elu, B, Cn, T, bx, tx, SN_THREADS); });
}
void emu_sn_linear_rows(const float* a, const float* wt, const float* bias, float* out, int R, int Kd, int N) {
  walk(sn_linear_rows_grid(R, N), [&](int bx, int by, int, int tx) { sn_linear_rows_thread(a, wt, bias, out, R, Kd, N, bx, by, tx, SN_THREADS); });
}
void emu_sn_lstm_step(const float* xg, const float* whh_t4, float* hseq, float* c, int t, int B, int H) {
  walk(sn_lstm_step_grid(B, H), [&](int bx, int by, int, int tx) { sn_lstm_step_thread(xg, whh_t4, hseq, c, t, B, H, bx, by, tx, SN_THREADS); });
}

void emu_sn_pack_conv_weight(const float* w, float* wp, int Co, int Ci, int K, int Cop, int transposed) {
  walk(sn_linear_grid_1d((long long)Ci * K * Cop),
       [&](int bx, int, int, int tx) { sn_pack_conv_weight_thread(w, wp, Co, Ci, K, Cop, transposed, bx, tx, SN_THREADS); });
}
void emu_sn_conv1d_packed(const pt_sn_conv_t* p) {
  if (sn_conv1d_packed_ct(*p) == 16)
    walk(sn_conv1d_packed_grid<16>(*p), [&](int bx, int by, int bz, int tx) { sn_conv1d_packed_thread<16>(*p, bx, by, bz, tx, SN_THREADS); });
  else
    walk(sn_conv1d_packed_grid<8>(*p), [&](int bx, int by, int bz, int tx) { sn_conv1d_packed_thread<8>(*p, bx, by, bz, tx, SN_THREADS); });
}
void emu_sn_conv_transpose1d_packed(const pt_sn_conv_t* p) {
  walk(sn_convtr_packed_grid(*p), [&](int bx, int by, int bz, int tx) { sn_convtr_packed_thread(*p, bx, by, bz, tx, SN_THREADS); });
}
// the cooperative kernel sn_lstm_seq_kernel of seanet.cu with its barriers turned into loop boundaries: per step every block
// stages + computes (blocks of one step are independent), "shared memory" is a per-block host array that lives across steps
void emu_sn_lstm_seq(const float* xg, const float* whh_t4, float* hseq, float* c, int T, int B, int H) {
  const int blocks = H / SN_PU, ntx = 32 * SN_PU;
  const size_t per_block = sn_lstm_seq_smem_floats(H);
  std::vector<float> smem(per_block * blocks, NAN);
  for (int bx = 0; bx < blocks; ++bx)
    for (int tx = 0; tx < ntx; ++tx) sn_lstm_seq_load_w(whh_t4, smem.data() + per_block * bx, H, bx, tx, ntx);
  for (int t = 0; t < T; ++t)
    for (int bx = 0; bx < blocks; ++bx) {
      float* wsm = smem.data() + per_block * bx;
      float* hs = wsm + (size_t)SN_PU * H * 4;
      for (int b0 = 0; b0 < B; b0 += 32) {
        const int nb = B - b0 < 32 ? B - b0 : 32;
        if (t > 0)
          for (int tx = 0; tx < ntx; ++tx) sn_lstm_seq_stage(hseq, hs, t, b0, nb, B, H, tx, ntx);
        for (int tx = 0; tx < ntx; ++tx) sn_lstm_seq_compute(xg, wsm, hs, hseq, c, t, b0, nb, B, H, bx, tx);
      }
    }
}

}  // extern "C"
