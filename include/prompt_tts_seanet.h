/* prompt_tts_seanet.h -- C ABI of libpt_seanet.so: EnCodec's SEANet encoder / decoder layers on B200 (sm_100a), fp32.
 *
 * SURVEY 8f row 4 ("next"): the waveform <-> 128-d latent stacks on either side of the RVQ kernels of prompt_tts_b200.h.
 * The reference reaches them through the third-party `encodec` package:
 *     /root/reference/data_preparation/generate_code.py:13-15,48   EncodecModel.encodec_model_24khz().encode(wav)
 *     /root/reference/decode_codec.py:8-9,16                        model.decode([(codes, None)])
 * encodec 0.1.1 is not vendored; its modules are restated by transformers.models.encodec.modeling_encodec (cited as ME:<line>,
 * transformers 5.5.0) -- each entry point names the module method it replaces.
 *
 * Conventions (same as prompt_tts_b200.h): plain pointers and sizes, every pointer is DEVICE memory, `stream` is a cudaStream_t,
 * returns 0 or a negative code (text from pt_sn_last_error()), never synchronises, owns no device memory.
 * Tensors at this boundary keep the reference's layout: activations [B, C, T] fp32 (T contiguous), Conv1d weights [Co, Ci, K],
 * ConvTranspose1d weights [Ci, Co, K].
 */
#ifndef PROMPT_TTS_SEANET_H
#define PROMPT_TTS_SEANET_H
#ifdef __cplusplus
extern "C" {
#endif

int pt_sn_version(void);
const char* pt_sn_last_error(void);
unsigned long long pt_sn_launch_count(void); /* kernels launched by this library since load */

/* Weight-norm fold, dim = 0 (nn.utils.parametrizations.weight_norm as applied at ME:103-108 / ME:171-176):
 * w[r, :] = g[r] * v[r, :] / |v[r, :]|_2 ; rows = size of the first axis, cols = product of the others. */
int pt_sn_weight_norm_fold(const float* v, const float* g, float* w, int rows, int cols, void* stream);

typedef struct {
  const float* x;     /* [B, Ci, Lin] */
  const float* w;     /* conv: [Co, Ci, K]; transposed conv: [Ci, Co, K] (weight norm already folded) */
  const float* bias;  /* [Co] or NULL */
  const float* res;   /* [B, Co, Lout] added to the result (the block's shortcut), or NULL */
  float* y;           /* [B, Co, Lout] result, or NULL */
  float* y_elu;       /* [B, Co, Lout] ELU(result), or NULL: the activation of the NEXT layer is applied by the producer */
  int B, Ci, Co, Lin, Lout, K, stride, dil;
  int pad_left;       /* conv: samples of padding in front of x[.., 0]; transposed conv: samples trimmed from the front */
  int reflect;        /* conv: 1 = reflect padding (F.pad 'reflect'), 0 = zeros */
  int Co_pad;         /* packed entry points only: row length of the packed weights (Co rounded up to a multiple of 8 or 16) */
} pt_sn_conv_t;

/* EncodecConv1d.forward (ME:150-170; encodec modules/conv.py SConv1d): y[b, co, t] = bias[co] +
 * sum_{ci, k} w[co, ci, k] * xpad[b, ci, t * stride + k * dil - pad_left] (+ res).  Padding is done by index arithmetic; the
 * caller chooses pad_left (all of (K-1)*dil+1-stride when causal) and Lout = ceil(Lin / stride) -- the extra right padding of
 * ME:125-133 is whatever the last window reaches past the end.  Error if a reflected index would leave the signal. */
int pt_sn_conv1d(const pt_sn_conv_t* p, void* stream);

/* EncodecConvTranspose1d.forward (ME:183-208; SConvTranspose1d): the full transposed convolution of length (Lin-1)*stride + K,
 * of which [pad_left, pad_left + Lout) is produced.  dil must be 1; `reflect` is ignored. */
int pt_sn_conv_transpose1d(const pt_sn_conv_t* p, void* stream);

/* Packed weights for the fast kernels: wp[(ci * K + k) * Co_pad + co] = w[co, ci, k] (conv) or w[ci, co, k] (transposed = 1), zero
 * for co >= Co.  Co_pad must be a multiple of 8.  A thread's 8 / 16 output channels are then two / four 16-byte loads. */
int pt_sn_pack_conv_weight(const float* w, float* wp, int Co, int Ci, int K, int Co_pad, int transposed, void* stream);
/* Same contracts as pt_sn_conv1d / pt_sn_conv_transpose1d with p->w = packed weights and p->Co_pad set: 16 output channels per
 * thread when Co_pad % 16 == 0, else 8; threads whose windows lie inside the signal run a loop without index checks. */
int pt_sn_conv1d_packed(const pt_sn_conv_t* p, void* stream);
int pt_sn_conv_transpose1d_packed(const pt_sn_conv_t* p, void* stream);

/* LSTM (EncodecLSTM.forward ME:219-223 = nn.LSTM, gate order i, f, g, o, zero initial state, + skip connection).
 * Packed weights: wt4[k][j][q] = W[q * H + j][k] for W = weight_ih / weight_hh [4H, H]; bias4[j][q] = b_ih[q*H+j] + b_hh[q*H+j]. */
int pt_sn_lstm_pack(const float* w, float* wt4, int H, void* stream);
int pt_sn_lstm_pack_bias(const float* b_ih, const float* b_hh, float* bias4, int H, void* stream);
/* [B, C, T] -> [T, B, C] */
int pt_sn_ncl_to_tbc(const float* x, float* out, int B, int Cn, int T, void* stream);
/* out[r, n] = bias[n] + sum_k a[r, k] * wt[k, n]   (a [R, Kd] row-major, wt [Kd, N], N % 4 == 0): the input projection of all steps */
int pt_sn_linear_rows(const float* a, const float* wt, const float* bias, float* out, int R, int Kd, int N, void* stream);
/* One time step for all B sequences: gates = xg[t] + h[t-1] * W_hh (packed), c and hseq[t] updated.  xg [T, B, H, 4],
 * hseq [T, B, H], c [B, H] (read only when t > 0).  Steps must be launched in order on one stream. */
int pt_sn_lstm_step(const float* xg, const float* whh_t4, float* hseq, float* c, int t, int B, int H, void* stream);
/* All T steps of one layer in ONE cooperative launch: H / 4 blocks, each keeps the recurrent weights of its 4 hidden units in shared
 * memory for the whole sequence, stages h[t-1] through shared memory and meets the other blocks at a grid-wide barrier per step.
 * Same arguments and results as T calls of pt_sn_lstm_step.  Returns -3 (nothing launched) if the device cannot keep H / 4 blocks
 * co-resident or has no cooperative launch: the caller then uses pt_sn_lstm_step. */
int pt_sn_lstm_seq(const float* xg, const float* whh_t4, float* hseq, float* c, int T, int B, int H, void* stream);
/* y[b, c, t] = hseq[t, b, c] + x[b, c, t] (the skip), written raw and / or through ELU */
int pt_sn_tbc_add_to_ncl(const float* hseq, const float* x, float* y, float* y_elu, int B, int Cn, int T, void* stream);

#ifdef __cplusplus
}
#endif
#endif
