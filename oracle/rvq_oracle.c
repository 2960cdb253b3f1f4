/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the EnCodec RVQ the reference calls.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (prompt_tts_b200) never does.
 *
 * Restates (the algorithm lives in third-party `encodec ^0.1.1`, reference pyproject.toml:11, not vendored; reached from
 * /root/reference/data_preparation/generate_code.py:48 and /root/reference/decode_codec.py:16; the identical formula is
 * in transformers 5.5 modeling_encodec.py:364-369,424-447, which is what tests/golden/rvq_*.npz were generated with):
 *   encode: for q in 0..Q-1: dist_j = -(|r|^2 - 2 r.e_j + |e_j|^2); idx_q = argmax_j dist_j; r -= e[idx_q]
 *   decode: latent = sum_q e_q[codes_q]   (q ascending, fp32, starting from 0)
 * Summation order is fixed here (the library GEMM the reference uses leaves it unspecified): every dot product and
 * squared norm is accumulated in ascending d with fmaf; the first maximal index wins ties.
 * Parity pin: tests/test_oracle.py checks this file against the golden vectors produced by the transformers
 * implementation (grid-valued data: must be identical; Gaussian data: identical wherever the fp64 margin is not tiny).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* latents [B, D, T] fp32, codebooks [Q, K, D] fp32 -> codes [B, Q, T] int64 */
void rvq_oracle_encode(const float* lat, const float* cb, int64_t* codes, int B, int D, int T, int Q, int K) {
  float* r = (float*)malloc(sizeof(float) * D);
  float* ee = (float*)malloc(sizeof(float) * (size_t)Q * K);
  for (long i = 0; i < (long)Q * K; ++i) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(cb[i * D + d], cb[i * D + d], s);
    ee[i] = s;
  }
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      for (int d = 0; d < D; ++d) r[d] = lat[((long)b * D + d) * T + t];
      for (int q = 0; q < Q; ++q) {
        float xx = 0.f;
        for (int d = 0; d < D; ++d) xx = fmaf(r[d], r[d], xx);
        float best = -INFINITY;
        int bi = 0;
        const float* cq = cb + (long)q * K * D;
        for (int j = 0; j < K; ++j) {
          float dot = 0.f;
          for (int d = 0; d < D; ++d) dot = fmaf(r[d], cq[(long)j * D + d], dot);
          volatile float a = 2.f * dot;
          volatile float s1 = xx - a;
          volatile float s2 = s1 + ee[(long)q * K + j];
          float dist = -s2;
          if (dist > best) {
            best = dist;
            bi = j;
          }
        }
        codes[((long)b * Q + q) * T + t] = bi;
        for (int d = 0; d < D; ++d) r[d] = r[d] - cq[(long)bi * D + d];
      }
    }
  free(r);
  free(ee);
}

/* codes [B, Q, T] int64 -> latents [B, D, T] fp32 */
void rvq_oracle_decode(const int64_t* codes, const float* cb, float* lat, int B, int D, int T, int Q, int K) {
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int d = 0; d < D; ++d) {
        volatile float acc = 0.f;
        for (int q = 0; q < Q; ++q) acc = acc + cb[((long)q * K + codes[((long)b * Q + q) * T + t]) * D + d];
        lat[((long)b * D + d) * T + t] = acc;
      }
}

/* x0 = (codes/1023 - 0.5)/0.5 (tts/dataloader.py:64,77,168-170), each step rounded to fp32 */
void codes_affine_oracle(const int64_t* codes, float* x0, long n) {
  for (long i = 0; i < n; ++i) {
    volatile float u = (float)codes[i] / 1023.f;
    volatile float v = u - 0.5f;
    x0[i] = v / 0.5f;
  }
}
