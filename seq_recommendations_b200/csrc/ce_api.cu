// C-ABI entry points of the logits + softmax cross-entropy path (K5/K6) and library sanity calls.
// These are the exact-fp32 SIMT kernels (ce_simt.cu); the tcgen05 kernels have their own entry points in ce_tc.cu.
#include "common.cuh"

int ce_forward_simt(const float* hout, const float* hscale, const float* W_out, const float* b_out,
                    const int32_t* tgt, float* ws_m, float* ws_s, float* zy, int64_t n_tokens, int H, int V,
                    int v_begin, int v_end, int ldw, int splits, cudaStream_t st);
int ce_backward_simt(const float* hout, const float* hscale, const float* W_out, const float* b_out,
                     const int32_t* tgt, const float* m, const float* s, const float* coef, const float* inv_nvalid,
                     float* dh, float* dW_out, float* db_out, int64_t n_tokens, int H, int v_begin, int v_end,
                     int ldw, int accumulate_dh, cudaStream_t st);

#include <atomic>
static std::atomic<int> g_launches{0};
void seqrec_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" int seqrec_abi_version(void) { return 1; }

extern "C" int seqrec_launch_count(int reset) {
  const int v = g_launches.load(std::memory_order_relaxed);
  if (reset) g_launches.store(0, std::memory_order_relaxed);
  return v;
}

extern "C" int seqrec_device_cc(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -(int)e;
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return -(int)e;
  return p.major * 10 + p.minor;
}

extern "C" int seqrec_ce_forward(const float* hout, const float* hscale, const float* W_out, const float* b_out,
                                 const int32_t* tgt, float* ws_m, float* ws_s, float* zy, int64_t n_tokens, int H,
                                 int V, int v_begin, int v_end, int ldw, int splits, void* stream) {
  SEQREC_ARG(n_tokens > 0 && H > 0 && V > 0, 1);
  SEQREC_ARG(v_begin >= 0 && v_end <= V && v_begin < v_end && ldw >= v_end && splits >= 1, 2);
  return ce_forward_simt(hout, hscale, W_out, b_out, tgt, ws_m, ws_s, zy, n_tokens, H, V, v_begin, v_end, ldw, splits,
                         as_stream(stream));
}

extern "C" int seqrec_ce_backward(const float* hout, const float* hscale, const float* W_out, const float* b_out,
                                  const int32_t* tgt, const float* m, const float* s, const float* coef,
                                  const float* inv_nvalid, float* dh, float* dW_out, float* db_out, int64_t n_tokens,
                                  int H, int V, int v_begin, int v_end, int ldw, int accumulate_dh, void* stream) {
  SEQREC_ARG(n_tokens > 0 && H > 0 && V > 0, 1);
  SEQREC_ARG(v_begin >= 0 && v_end <= V && v_begin < v_end, 2);
  return ce_backward_simt(hout, hscale, W_out, b_out, tgt, m, s, coef, inv_nvalid, dh, dW_out, db_out, n_tokens, H,
                          v_begin, v_end, ldw, accumulate_dh, as_stream(stream));
}
