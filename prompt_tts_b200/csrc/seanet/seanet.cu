// seanet.cu -- libpt_seanet.so: EnCodec SEANet encoder / decoder layers for sm_100a (include/prompt_tts_seanet.h).
//
// fp32 on the FMA pipe, the reference's [B, C, T] layout: these stacks feed the RVQ quantiser, whose codes flip on rounding noise, so
// this path keeps the reference's precision.  The kernel bodies are in seanet_core.h (shared with the host-side index checker of
// the CPU test tier); this file is the __global__ wrappers, argument validation and launches.  Default path: packed-weight
// convolutions + one cooperative launch per LSTM layer; the first-draft kernels stay as A/B partners and as the LSTM fallback.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "seanet_core.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-2, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

dim3 to_dim3(const sn_grid& g) { return dim3(g.x, g.y, g.z); }

__global__ void __launch_bounds__(SN_THREADS) sn_conv1d_kernel(const pt_sn_conv_t p) {
  sn_conv1d_thread(p, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_convtr_kernel(const pt_sn_conv_t p) {
  sn_convtr_thread(p, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_weight_norm_kernel(const float* v, const float* g, float* w, int rows, int cols) {
  sn_weight_norm_thread(v, g, w, rows, cols, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_lstm_pack_kernel(const float* w, float* wt4, int H) {
  sn_lstm_pack_thread(w, wt4, H, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_lstm_pack_bias_kernel(const float* b_ih, const float* b_hh, float* bias4, int H) {
  sn_lstm_pack_bias_thread(b_ih, b_hh, bias4, H, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_ncl_to_tbc_kernel(const float* x, float* out, int B, int Cn, int T) {
  sn_ncl_to_tbc_thread(x, out, B, Cn, T, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS)
    sn_tbc_add_to_ncl_kernel(const float* hseq, const float* x, float* y, float* y_elu, int B, int Cn, int T) {
  sn_tbc_add_to_ncl_thread(hseq, x, y, y_elu, B, Cn, T, blockIdx.x, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS)
    sn_linear_rows_kernel(const float* a, const float* wt, const float* bias, float* out, int R, int Kd, int N) {
  sn_linear_rows_thread(a, wt, bias, out, R, Kd, N, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS)
    sn_lstm_step_kernel(const float* xg, const float* whh_t4, float* hseq, float* c, int t, int B, int H) {
  sn_lstm_step_thread(xg, whh_t4, hseq, c, t, B, H, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

template <int CT>
__global__ void __launch_bounds__(SN_THREADS) sn_conv1d_packed_kernel(const pt_sn_conv_t p) {
  sn_conv1d_packed_thread<CT>(p, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS) sn_convtr_packed_kernel(const pt_sn_conv_t p) {
  sn_convtr_packed_thread(p, blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x, blockDim.x);
}
__global__ void __launch_bounds__(SN_THREADS)
    sn_pack_conv_weight_kernel(const float* w, float* wp, int Co, int Ci, int K, int Cop, int transposed) {
  sn_pack_conv_weight_thread(w, wp, Co, Ci, K, Cop, transposed, blockIdx.x, threadIdx.x, blockDim.x);
}

// One layer of the LSTM over all T steps: H / SN_PU co-resident blocks (cooperative launch), one grid-wide barrier per step.
__global__ void __launch_bounds__(32 * SN_PU)
    sn_lstm_seq_kernel(const float* xg, const float* whh_t4, float* hseq, float* c, int T, int B, int H) {
  extern __shared__ __align__(16) float sn_smem[];
  float* wsm = sn_smem;                           // [SN_PU][H][4]
  float* hs = sn_smem + (size_t)SN_PU * H * 4;    // [32][H + 4]
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  sn_lstm_seq_load_w(whh_t4, wsm, H, blockIdx.x, threadIdx.x, blockDim.x);
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int nb = B - b0 < 32 ? B - b0 : 32;
      if (t > 0) sn_lstm_seq_stage(hseq, hs, t, b0, nb, B, H, threadIdx.x, blockDim.x);
      __syncthreads();
      sn_lstm_seq_compute(xg, wsm, hs, hseq, c, t, b0, nb, B, H, blockIdx.x, threadIdx.x);
      __syncthreads();  // hs is overwritten by the next chunk / step
    }
    if (t + 1 < T) grid.sync();  // h[t] of every block is visible before anyone stages it
  }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_conv_common(const pt_sn_conv_t* p, const char* what) {
  if (!p) return fail(-1, "%s: null descriptor", what);
  if (!p->x || !p->w) return fail(-1, "%s: x and w must be set", what);
  if (!p->y && !p->y_elu) return fail(-1, "%s: neither y nor y_elu is set", what);
  if (p->B <= 0 || p->Ci <= 0 || p->Co <= 0 || p->Lin <= 0 || p->Lout <= 0 || p->K <= 0 || p->stride <= 0 || p->dil <= 0 ||
      p->pad_left < 0)
    return fail(-1, "%s: bad sizes B=%d Ci=%d Co=%d Lin=%d Lout=%d K=%d stride=%d dil=%d pad_left=%d", what, p->B, p->Ci, p->Co, p->Lin,
                p->Lout, p->K, p->stride, p->dil, p->pad_left);
  if (p->B > 65535) return fail(-1, "%s: B=%d exceeds the grid's z extent", what, p->B);
  // sample positions are 32-bit in the kernels: (Lout + one block tile) * stride + K * dil must fit
  const long long reach = ((long long)p->Lout + SN_THREADS * SN_TT) * p->stride + (long long)p->K * p->dil + p->pad_left;
  if (reach > 0x7fffffffLL || (long long)p->Lin > 0x7fffffffLL - SN_THREADS * SN_TT)
    return fail(-1, "%s: sequence too long for 32-bit sample positions (Lin=%d Lout=%d stride=%d)", what, p->Lin, p->Lout, p->stride);
  return 0;
}

}  // namespace

extern "C" {

int pt_sn_version(void) { return 1; }
const char* pt_sn_last_error(void) { return g_err; }
unsigned long long pt_sn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int pt_sn_weight_norm_fold(const float* v, const float* g, float* w, int rows, int cols, void* stream) {
  if (!v || !g || !w || rows <= 0 || cols <= 0) return fail(-1, "weight_norm_fold: bad arguments");
  sn_weight_norm_kernel<<<to_dim3(sn_linear_grid_1d(rows)), SN_THREADS, 0, (cudaStream_t)stream>>>(v, g, w, rows, cols);
  return launched("weight_norm_fold");
}

int pt_sn_conv1d(const pt_sn_conv_t* p, void* stream) {
  if (int r = check_conv_common(p, "conv1d")) return r;
  const long long last = (long long)(p->Lout - 1) * p->stride + (long long)(p->K - 1) * p->dil - p->pad_left;  // last sample touched
  if (p->reflect) {
    if (p->pad_left > p->Lin - 1 || last > 2LL * (p->Lin - 1))
      return fail(-1, "conv1d: input of length %d is shorter than its reflect padding (%d in front, %lld behind)", p->Lin, p->pad_left,
                  last - (p->Lin - 1));
  }
  const sn_grid g = sn_conv1d_grid(*p);
  if (g.y > 65535) return fail(-1, "conv1d: Co=%d too large", p->Co);
  sn_conv1d_kernel<<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return launched("conv1d");
}

int pt_sn_conv_transpose1d(const pt_sn_conv_t* p, void* stream) {
  if (int r = check_conv_common(p, "conv_transpose1d")) return r;
  if (p->dil != 1) return fail(-1, "conv_transpose1d: dilation %d is not supported", p->dil);
  const long long full = (long long)(p->Lin - 1) * p->stride + p->K;
  if ((long long)p->pad_left + p->Lout > full)
    return fail(-1, "conv_transpose1d: window [%d, %lld) exceeds the full output length %lld", p->pad_left, (long long)p->pad_left + p->Lout,
                full);
  const sn_grid g = sn_convtr_grid(*p);
  if (g.y > 65535) return fail(-1, "conv_transpose1d: Co * stride too large");
  sn_convtr_kernel<<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return launched("conv_transpose1d");
}

int pt_sn_pack_conv_weight(const float* w, float* wp, int Co, int Ci, int K, int Co_pad, int transposed, void* stream) {
  if (!w || !wp || Co <= 0 || Ci <= 0 || K <= 0 || Co_pad < Co || Co_pad % 8) return fail(-1, "pack_conv_weight: bad arguments");
  sn_pack_conv_weight_kernel<<<to_dim3(sn_linear_grid_1d((long long)Ci * K * Co_pad)), SN_THREADS, 0, (cudaStream_t)stream>>>(
      w, wp, Co, Ci, K, Co_pad, transposed);
  return launched("pack_conv_weight");
}

static int check_packed(const pt_sn_conv_t* p, const char* what) {
  if (p->Co_pad < p->Co || p->Co_pad % 8) return fail(-1, "%s: Co_pad=%d must be a multiple of 8 that is >= Co=%d", what, p->Co_pad, p->Co);
  if (!aligned16(p->w)) return fail(-1, "%s: packed weights must be 16-byte aligned", what);
  return 0;
}

int pt_sn_conv1d_packed(const pt_sn_conv_t* p, void* stream) {
  if (int r = check_conv_common(p, "conv1d_packed")) return r;
  if (int r = check_packed(p, "conv1d_packed")) return r;
  const long long last = (long long)(p->Lout - 1) * p->stride + (long long)(p->K - 1) * p->dil - p->pad_left;
  if (p->reflect && (p->pad_left > p->Lin - 1 || last > 2LL * (p->Lin - 1)))
    return fail(-1, "conv1d_packed: input of length %d is shorter than its reflect padding (%d in front, %lld behind)", p->Lin,
                p->pad_left, last - (p->Lin - 1));
  if (sn_conv1d_packed_ct(*p) == 16) {
    const sn_grid g = sn_conv1d_packed_grid<16>(*p);
    if (g.y > 65535) return fail(-1, "conv1d_packed: Co=%d too large", p->Co);
    sn_conv1d_packed_kernel<16><<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(*p);
  } else {
    const sn_grid g = sn_conv1d_packed_grid<8>(*p);
    if (g.y > 65535) return fail(-1, "conv1d_packed: Co=%d too large", p->Co);
    sn_conv1d_packed_kernel<8><<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(*p);
  }
  return launched("conv1d_packed");
}

int pt_sn_conv_transpose1d_packed(const pt_sn_conv_t* p, void* stream) {
  if (int r = check_conv_common(p, "conv_transpose1d_packed")) return r;
  if (int r = check_packed(p, "conv_transpose1d_packed")) return r;
  if (p->dil != 1) return fail(-1, "conv_transpose1d_packed: dilation %d is not supported", p->dil);
  const long long full = (long long)(p->Lin - 1) * p->stride + p->K;
  if ((long long)p->pad_left + p->Lout > full)
    return fail(-1, "conv_transpose1d_packed: window [%d, %lld) exceeds the full output length %lld", p->pad_left,
                (long long)p->pad_left + p->Lout, full);
  const sn_grid g = sn_convtr_packed_grid(*p);
  if (g.y > 65535) return fail(-1, "conv_transpose1d_packed: Co * stride too large");
  sn_convtr_packed_kernel<<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(*p);
  return launched("conv_transpose1d_packed");
}

int pt_sn_lstm_seq(const float* xg, const float* whh_t4, float* hseq, float* c, int T, int B, int H, void* stream) {
  if (!xg || !whh_t4 || !hseq || !c || T <= 0 || B <= 0 || H <= 0) return fail(-1, "lstm_seq: bad arguments");
  if (H % SN_PU || H % 4) return fail(-1, "lstm_seq: H=%d must be a multiple of %d", H, SN_PU);
  if (!aligned16(xg) || !aligned16(whh_t4) || !aligned16(hseq)) return fail(-1, "lstm_seq: pointers must be 16-byte aligned");
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(-2, "lstm_seq: cudaGetDevice failed");
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop) return fail(-3, "lstm_seq: the device has no cooperative launch");
  const size_t smem = sn_lstm_seq_smem_floats(H) * sizeof(float);
  if (smem > 227 * 1024) return fail(-3, "lstm_seq: H=%d needs %zu bytes of shared memory", H, smem);
  cudaError_t e = cudaFuncSetAttribute(sn_lstm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(-3, "lstm_seq: shared memory opt-in failed: %s", cudaGetErrorString(e));
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sn_lstm_seq_kernel, 32 * SN_PU, smem);
  const int blocks = H / SN_PU;
  if (e != cudaSuccess || per_sm * sms < blocks)
    return fail(-3, "lstm_seq: %d blocks cannot be co-resident (%d per SM x %d SMs)", blocks, per_sm, sms);
  void* args[] = {(void*)&xg, (void*)&whh_t4, (void*)&hseq, (void*)&c, (void*)&T, (void*)&B, (void*)&H};
  e = cudaLaunchCooperativeKernel((const void*)sn_lstm_seq_kernel, dim3(blocks), dim3(32 * SN_PU), args, smem, (cudaStream_t)stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(-2, "lstm_seq: %s", cudaGetErrorString(e));
  return 0;
}

int pt_sn_lstm_pack(const float* w, float* wt4, int H, void* stream) {
  if (!w || !wt4 || H <= 0) return fail(-1, "lstm_pack: bad arguments");
  sn_lstm_pack_kernel<<<to_dim3(sn_linear_grid_1d(4LL * H * H)), SN_THREADS, 0, (cudaStream_t)stream>>>(w, wt4, H);
  return launched("lstm_pack");
}

int pt_sn_lstm_pack_bias(const float* b_ih, const float* b_hh, float* bias4, int H, void* stream) {
  if (!b_ih || !b_hh || !bias4 || H <= 0) return fail(-1, "lstm_pack_bias: bad arguments");
  sn_lstm_pack_bias_kernel<<<to_dim3(sn_linear_grid_1d(4LL * H)), SN_THREADS, 0, (cudaStream_t)stream>>>(b_ih, b_hh, bias4, H);
  return launched("lstm_pack_bias");
}

int pt_sn_ncl_to_tbc(const float* x, float* out, int B, int Cn, int T, void* stream) {
  if (!x || !out || B <= 0 || Cn <= 0 || T <= 0) return fail(-1, "ncl_to_tbc: bad arguments");
  sn_ncl_to_tbc_kernel<<<to_dim3(sn_linear_grid_1d((long long)B * Cn * T)), SN_THREADS, 0, (cudaStream_t)stream>>>(x, out, B, Cn, T);
  return launched("ncl_to_tbc");
}

int pt_sn_tbc_add_to_ncl(const float* hseq, const float* x, float* y, float* y_elu, int B, int Cn, int T, void* stream) {
  if (!hseq || !x || (!y && !y_elu) || B <= 0 || Cn <= 0 || T <= 0) return fail(-1, "tbc_add_to_ncl: bad arguments");
  sn_tbc_add_to_ncl_kernel<<<to_dim3(sn_linear_grid_1d((long long)B * Cn * T)), SN_THREADS, 0, (cudaStream_t)stream>>>(hseq, x, y, y_elu, B,
                                                                                                                   Cn, T);
  return launched("tbc_add_to_ncl");
}

int pt_sn_linear_rows(const float* a, const float* wt, const float* bias, float* out, int R, int Kd, int N, void* stream) {
  if (!a || !wt || !out || R <= 0 || Kd <= 0 || N <= 0) return fail(-1, "linear_rows: bad arguments");
  if (Kd % 4 || N % 4) return fail(-1, "linear_rows: Kd=%d and N=%d must be multiples of 4", Kd, N);
  if (!aligned16(a) || !aligned16(wt) || !aligned16(out) || (bias && !aligned16(bias)))
    return fail(-1, "linear_rows: pointers must be 16-byte aligned");
  const sn_grid g = sn_linear_rows_grid(R, N);
  if (g.y > 65535) {  // split the rows over several launches
    const int rows_per = 65535 * SN_LR;
    for (int r = 0; r < R; r += rows_per) {
      const int n = R - r < rows_per ? R - r : rows_per;
      if (int rc = pt_sn_linear_rows(a + (size_t)r * Kd, wt, bias, out + (size_t)r * N, n, Kd, N, stream)) return rc;
    }
    return 0;
  }
  sn_linear_rows_kernel<<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(a, wt, bias, out, R, Kd, N);
  return launched("linear_rows");
}

int pt_sn_lstm_step(const float* xg, const float* whh_t4, float* hseq, float* c, int t, int B, int H, void* stream) {
  if (!xg || !whh_t4 || !hseq || !c || t < 0 || B <= 0 || H <= 0) return fail(-1, "lstm_step: bad arguments");
  if (H % 4) return fail(-1, "lstm_step: H=%d must be a multiple of 4", H);
  if (!aligned16(xg) || !aligned16(whh_t4) || !aligned16(hseq)) return fail(-1, "lstm_step: pointers must be 16-byte aligned");
  const sn_grid g = sn_lstm_step_grid(B, H);
  sn_lstm_step_kernel<<<to_dim3(g), SN_THREADS, 0, (cudaStream_t)stream>>>(xg, whh_t4, hseq, c, t, B, H);
  return launched("lstm_step");
}

}  // extern "C"
