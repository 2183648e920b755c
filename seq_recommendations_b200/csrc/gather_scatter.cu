// K1 row gather, K7 warp-aggregated scatter-add, and the (B,T)->(T,B) batch formatter.
//
// Replaces the dense one-hot (N,V).(V,G*H) product Theano runs for `LSTM(...)(masked one-hot)` (model.py:360-364)
// and the dense one_hot^T . dxp gradient.  Both are HBM-bound; the roofline figures in DESIGN.md are
// N*(4 + 2*G*H*4) B for the gather and N*(4 + 3*G*H*4) B for the scatter-add.
#include "common.cuh"

// ----------------------------------------------------------------------------------------------------------------
// (B,T) -> (T,B) ids/targets/mask; pad = negative id.  Also counts valid tokens.
__global__ void format_batch_kernel(const int32_t* __restrict__ ids_bt, const int32_t* __restrict__ tgt_bt,
                                    int32_t* __restrict__ ids_tb, int32_t* __restrict__ tgt_tb,
                                    uint8_t* __restrict__ mask_tb, int32_t* __restrict__ n_valid, int B, int T,
                                    int n_in, int n_items, int32_t* __restrict__ err) {
  __shared__ int32_t tile_i[32][33];
  __shared__ int32_t tile_t[32][33];
  const int b0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  int cnt = 0;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int b = b0 + r, t = t0 + threadIdx.x;
    int32_t vi = -1, vt = -1;
    if (b < B && t < T) {
      vi = ids_bt[(int64_t)b * T + t];
      vt = tgt_bt ? tgt_bt[(int64_t)b * T + t] : 0;
    }
    tile_i[r][threadIdx.x] = vi;
    tile_t[r][threadIdx.x] = vt;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int t = t0 + r, b = b0 + threadIdx.x;
    if (t < T && b < B) {
      int32_t vi = tile_i[threadIdx.x][r];
      int32_t vt = tile_t[threadIdx.x][r];
      // range check (the reference raises IndexError in np_utils.to_categorical for such an id): an id >= n_in or a
      // target >= n_items never reaches a kernel that would index a table with it -- the token becomes a pad and the
      // error flag is raised (bit 0: input id, bit 1: target), which the host turns into an exception
      if (vi >= n_in) { vi = -1; if (err) atomicOr(err, 1); }
      bool ok = vi >= 0;
      if (ok && tgt_tb && (vt >= n_items || vt < 0)) { ok = false; vi = -1; if (err) atomicOr(err, 2); }
      ids_tb[(int64_t)t * B + b] = vi;
      if (tgt_tb) tgt_tb[(int64_t)t * B + b] = ok ? vt : -1;
      mask_tb[(int64_t)t * B + b] = ok ? 1 : 0;
      cnt += ok ? 1 : 0;
    }
  }
  cnt = (int)warp_sum((float)cnt);
  if (threadIdx.x == 0 && cnt) atomicAdd(n_valid, cnt);
}

extern "C" int seqrec_format_batch(const int32_t* ids_bt, const int32_t* tgt_bt, int32_t* ids_tb, int32_t* tgt_tb,
                                   uint8_t* mask_tb, int32_t* n_valid, int B, int T, int n_in, int n_items,
                                   int32_t* err, void* stream) {
  SEQREC_ARG(B > 0 && T > 0 && n_in > 0 && n_items > 0, 1);
  dim3 grid(ceil_div(T, 32), ceil_div(B, 32)), block(32, 8);
  format_batch_kernel<<<grid, block, 0, as_stream(stream)>>>(ids_bt, tgt_bt, ids_tb, tgt_tb, mask_tb, n_valid, B, T,
                                                             n_in, n_items, err);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Ragged sequences -> the (B,T) id / target batch of FullModelPreprocessor (preprocessor.py:67-94 with Keras'
// pad_sequences(padding='pre', truncating='pre')): sequence i = items[offs[i] .. offs[i+1]) of length L gives the
// L-1 pairs (s[j], s[j+1]); the LAST min(L-1, T) pairs fill the right end of row i, the rest is pad (-1).
__global__ void __launch_bounds__(256)
pad_sequences_kernel(const int32_t* __restrict__ items, const int64_t* __restrict__ offs, int32_t* __restrict__ ids_bt,
                     int32_t* __restrict__ tgt_bt, int64_t n_seqs, int T) {
  const int64_t total = n_seqs * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / T;
    const int t = (int)(i - b * T);
    const int64_t start = offs[b], len = offs[b + 1] - start;
    const int64_t pairs = len > 0 ? len - 1 : 0;
    const int64_t keep = pairs < T ? pairs : T;                // pairs that fit after left truncation
    const int64_t j = t - (T - keep);                          // index among the kept pairs
    int32_t vi = -1, vt = -1;
    if (j >= 0) {
      const int64_t src = start + (pairs - keep) + j;
      vi = items[src];
      vt = items[src + 1];
    }
    ids_bt[i] = vi;
    if (tgt_bt) tgt_bt[i] = vt;
  }
}

extern "C" int seqrec_pad_sequences(const int32_t* items, const int64_t* offsets, int32_t* ids_bt, int32_t* tgt_bt,
                                    int64_t n_seqs, int T, void* stream) {
  SEQREC_ARG(items && offsets && ids_bt && n_seqs > 0 && T > 0, 1);
  int64_t blocks = (n_seqs * T + 255) / 256;
  if (blocks > SEQREC_NUM_SMS * 16) blocks = SEQREC_NUM_SMS * 16;
  pad_sequences_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(items, offsets, ids_bt, tgt_bt, n_seqs, T);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// History features (datasets.py:97-113 build_xs, then FullModelPreprocessor's c = xs[:-1] left-padded / left-truncated,
// preprocessor.py:71,89,92): c[b,t,v] = how often (freq) / whether item v occurred among s[0..j] of sequence b, j being
// the input position that lands on column t; pad columns are 0.  Rows that the left truncation drops still count -- the
// reference builds xs on the whole sequence and truncates afterwards.  One thread per (sequence, item): it walks the
// sequence once with the running count in a register; consecutive threads own consecutive items, so each store of a
// warp is one contiguous run of the (T, V) row block.  HBM-bound: n_seqs * T * V * 4 bytes written.
// table: optional value transform indexed by the count (the drivers' np.log(x + 1), experiments_server.py:35-36,
// computed on the host in float64 and rounded once, like the reference's own down-cast at the Theano boundary).
// The inner loop is a handful of instructions per stored vector: the running counts and their (rarely changing) output
// values live in registers, a match of the walked item with one of the thread's VEC items is the rare branch, and the
// items are fetched four steps ahead.  (First version: 25 instructions per scalar store, issue-bound at 0.58 of the HBM
// peak by ncu; profiles/r2_summary.md.)
template <int VEC>
struct HistoryRow {
  int cnt[VEC];
  float val[VEC];
  bool bad;
  __device__ __forceinline__ float value(int c, int freq, const float* __restrict__ table, int table_len) const {
    const int e = freq ? c : (c > 0 ? 1 : 0);
    return table ? __ldg(table + (e < table_len ? e : table_len - 1)) : (float)e;
  }
  __device__ __forceinline__ void init(int freq, const float* __restrict__ table, int table_len) {
    bad = false;
#pragma unroll
    for (int k = 0; k < VEC; ++k) { cnt[k] = 0; val[k] = value(0, freq, table, table_len); }
  }
  __device__ __forceinline__ void see(int32_t it, int v0, int V, int freq, const float* __restrict__ table,
                                      int table_len) {
    bad |= (unsigned)it >= (unsigned)V;
    const unsigned d = (unsigned)(it - v0);
    if (d < (unsigned)VEC) {
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (d == (unsigned)k) { cnt[k] += 1; val[k] = value(cnt[k], freq, table, table_len); }
    }
  }
  __device__ __forceinline__ void store(float* o) const {
    if (VEC == 4) __stcs(reinterpret_cast<float4*>(o), make_float4(val[0], val[1 % VEC], val[2 % VEC], val[3 % VEC]));
    else __stcs(o, val[0]);
  }
};

template <int VEC>
__global__ void __launch_bounds__(256)
history_features_kernel(const int32_t* __restrict__ items, const int64_t* __restrict__ offs, float* __restrict__ out,
                        int64_t n_seqs, int T, int V, int freq, const float* __restrict__ table, int table_len,
                        int32_t* __restrict__ err) {
  const int units = V / VEC;                                   // threads per sequence
  const int64_t total = n_seqs * units;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / units;
    const int v0 = (int)(i - b * units) * VEC;
    const int64_t start = offs[b], len = offs[b + 1] - start;
    const int64_t rows = len > 0 ? len - 1 : 0;                // xs[:-1]: one row per input position
    const int keep = (int)(rows < T ? rows : T);
    const int64_t skip = rows - keep;
    const int32_t* __restrict__ ip = items + start;
    HistoryRow<VEC> h;
    h.init(freq, table, table_len);
    for (int64_t j = 0; j < skip; ++j) h.see(__ldg(ip + j), v0, V, freq, table, table_len);
    ip += skip;
    float* o = out + (b * T) * (int64_t)V + v0;
    const int pad = T - keep;
    for (int t = 0; t < pad; ++t, o += V) {
      if (VEC == 4) __stcs(reinterpret_cast<float4*>(o), make_float4(0.f, 0.f, 0.f, 0.f));
      else __stcs(o, 0.f);
    }
    int j = 0;
    for (; j + 4 <= keep; j += 4) {
      const int32_t i0 = __ldg(ip + j), i1 = __ldg(ip + j + 1), i2 = __ldg(ip + j + 2), i3 = __ldg(ip + j + 3);
      h.see(i0, v0, V, freq, table, table_len); h.store(o); o += V;
      h.see(i1, v0, V, freq, table, table_len); h.store(o); o += V;
      h.see(i2, v0, V, freq, table, table_len); h.store(o); o += V;
      h.see(i3, v0, V, freq, table, table_len); h.store(o); o += V;
    }
    for (; j < keep; ++j, o += V) {
      h.see(__ldg(ip + j), v0, V, freq, table, table_len);
      h.store(o);
    }
    // (the reference's xi[s] raises IndexError for s >= V and silently wraps a negative s: both are errors here)
    if (v0 == 0 && h.bad && err) atomicOr(err, 1);
  }
}

// Small catalogs (V <= 32: MSNBC's 17 page categories -- where the reference's history models actually run): a group of
// LPS lanes (the next power of two >= V) owns one sequence, lane v its item v; a CTA owns 8 * (32 / LPS) consecutive
// sequences.  The group fetches LPS items with one coalesced load and hands them round by shuffle; the walk is
// branch-free (with V this small SOME lane of the warp matches at every step, so a "rare" branch is taken every time and
// serialises the steps: 0.30 of the HBM peak), the value table sits in shared memory, and the unrolled steps of a chunk
// overlap.  (Staging the rows in shared memory and writing the CTA's contiguous region with float4 stores was measured
// SLOWER -- 0.27 ms against 0.20 ms at the MSNBC shape: the walk, not the 68-byte row stores, is what binds.)
constexpr int HIST_TAB = 512;                                  // value-table entries kept in shared memory

// MODE fixes the value transform at compile time (a run-time choice costs a dozen predicated instructions per step):
// 0 = presence 0/1, 1 = counts, 2 = table[count] from shared memory, 3 = anything else (run-time flags).
template <int LPS, int MODE>
__global__ void __launch_bounds__(256)
history_features_small_kernel(const int32_t* __restrict__ items, const int64_t* __restrict__ offs,
                              float* __restrict__ out, int64_t n_seqs, int T, int V, int freq,
                              const float* __restrict__ table, int table_len, int32_t* __restrict__ err) {
  // (the entries past the table repeat its last one: MODE 2 walks a pointer through the table and clamps it once per
  //  chunk of LPS steps instead of at every step)
  __shared__ float tab_s[HIST_TAB + 32];
  constexpr int GPW = 32 / LPS;                                // sequences per warp
  constexpr int SPC = 8 * GPW;                                 // sequences per CTA
  const float* tab = table;
  if ((MODE == 2 || MODE == 3) && table && table_len <= HIST_TAB) {
    for (int i = threadIdx.x; i < HIST_TAB + 32; i += 256)
      tab_s[i] = __ldg(table + (i < table_len ? i : table_len - 1));
    __syncthreads();
    tab = tab_s;
  }
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPS, v = lane % LPS;
  const int local = (threadIdx.x >> 5) * GPW + sub;            // sequence within the CTA
  const int64_t b_base = (int64_t)blockIdx.x * SPC;
  const int64_t b = b_base + local;
  const bool live = b < n_seqs;
  int64_t start = 0, len = 0;
  if (live) { start = offs[b]; len = offs[b + 1] - start; }
  const int64_t rows = len > 0 ? len - 1 : 0;
  const int keep = (int)(rows < T ? rows : T);
  const int64_t skip = rows - keep;
  const int pad = T - keep;
  const int32_t* __restrict__ ip = items + start;
  const bool writer = live && v < V;
  const int last = table_len - 1;
  int cnt = 0, bad = 0;
  for (int64_t j = 0; j < skip; ++j) {                         // (rare: rows dropped by the left truncation still count)
    const int32_t it = __ldg(ip + j);
    bad |= ((unsigned)it >= (unsigned)V) ? 1 : 0;
    cnt += (it == v) ? 1 : 0;
  }
  ip += skip;
  float* o = out + (b * T) * (int64_t)V + v;
  if (writer)
    for (int t = 0; t < pad; ++t) __stcs(o + (size_t)t * V, 0.f);
  o += (size_t)pad * V;
  int kmax = keep;
#pragma unroll
  for (int d = LPS; d < 32; d <<= 1) {
    const int other = __shfl_xor_sync(0xffffffffu, kmax, d);
    kmax = other > kmax ? other : kmax;
  }
  // MODE 0 carries the output value itself (0 until the first match), MODE 2 the shared-memory address of the table
  // entry of the running count (+4 bytes per match): a step is shuffle, compare, predicated update, load, store
  float seen = cnt > 0 ? 1.f : 0.f;
  const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(tab_s);
  uint32_t va = tab_base + 4u * (uint32_t)(cnt < HIST_TAB - 1 ? cnt : HIST_TAB - 1);
  for (int j0 = 0; j0 < kmax; j0 += LPS, o += (size_t)LPS * V) {
    const int left = keep - j0;                                // steps of this chunk that exist for this sequence
    // (steps past the end of the sequence carry the item -1, which matches no lane)
    const int32_t chunk = (v < left) ? __ldg(ip + j0 + v) : -1;
    bad |= (v < left && (unsigned)chunk >= (unsigned)V) ? 1 : 0;   // every item is checked once, by the lane that loaded it
    if (MODE == 2) va = va < tab_base + 4u * (HIST_TAB - 1) ? va : tab_base + 4u * (HIST_TAB - 1);
#pragma unroll
    for (int jj = 0; jj < LPS; ++jj) {
      const int32_t it = __shfl_sync(0xffffffffu, chunk, jj, LPS);
      const bool hit = it == v;
      float val;
      if (MODE == 0) {
        seen = hit ? 1.f : seen;
        val = seen;
      } else if (MODE == 2) {
        va += hit ? 4u : 0u;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(va));
      } else {
        cnt += hit ? 1 : 0;
        if (MODE == 1) val = (float)cnt;
        else {
          const int e = freq ? cnt : (cnt > 0 ? 1 : 0);
          val = tab ? tab[e < last ? e : last] : (float)e;
        }
      }
      if (writer && jj < left) __stcs(o + (size_t)jj * V, val);
    }
  }
  if (live && bad && err) atomicOr(err, 1);
}

template <int LPS, int MODE>
int launch_history_small_mode(const int32_t* items, const int64_t* offsets, float* c, int64_t n_seqs, int T, int V,
                              int freq, const float* table, int table_len, int32_t* err, cudaStream_t st) {
  constexpr int SPC = 8 * (32 / LPS);
  const int64_t ctas = (n_seqs + SPC - 1) / SPC;
  SEQREC_ARG(ctas <= 0x7fffffff, 3);
  history_features_small_kernel<LPS, MODE><<<(int)ctas, 256, 0, st>>>(items, offsets, c, n_seqs, T, V, freq, table,
                                                                      table_len, err);
  return 0;
}

template <int LPS>
int launch_history_small(const int32_t* items, const int64_t* offsets, float* c, int64_t n_seqs, int T, int V,
                         int freq, const float* table, int table_len, int32_t* err, cudaStream_t st) {
  if (!table && !freq)
    return launch_history_small_mode<LPS, 0>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
  if (!table)
    return launch_history_small_mode<LPS, 1>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
  if (freq && table_len <= HIST_TAB)
    return launch_history_small_mode<LPS, 2>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
  return launch_history_small_mode<LPS, 3>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
}

extern "C" int seqrec_history_features(const int32_t* items, const int64_t* offsets, float* c, int64_t n_seqs, int T,
                                       int V, int freq, const float* table, int table_len, int32_t* err,
                                       void* stream) {
  SEQREC_ARG(items && offsets && c && n_seqs > 0 && T > 0 && V > 0, 1);
  SEQREC_ARG(!table || table_len > 1, 2);
  if (V <= 32) {
    cudaStream_t st = as_stream(stream);
    int rc;
    if (V > 16) rc = launch_history_small<32>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
    else if (V > 8) rc = launch_history_small<16>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
    else if (V > 4) rc = launch_history_small<8>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
    else rc = launch_history_small<4>(items, offsets, c, n_seqs, T, V, freq, table, table_len, err, st);
    if (rc) return rc;
    SEQREC_CHECK_LAUNCH();
    return 0;
  }
  const bool vec = (V % 4 == 0) && (reinterpret_cast<uintptr_t>(c) % 16 == 0);
  int64_t blocks = (n_seqs * (vec ? V / 4 : V) + 255) / 256;
  if (blocks > SEQREC_NUM_SMS * 8) blocks = SEQREC_NUM_SMS * 8;
  if (vec)
    history_features_kernel<4><<<(int)blocks, 256, 0, as_stream(stream)>>>(items, offsets, c, n_seqs, T, V, freq, table,
                                                                           table_len, err);
  else
    history_features_kernel<1><<<(int)blocks, 256, 0, as_stream(stream)>>>(items, offsets, c, n_seqs, T, V, freq, table,
                                                                           table_len, err);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// K1: xp[n,:] = (mask ? scale*W_in[id,:] : 0) + b.   One thread per float4 of the output, consecutive threads on
// consecutive 16-byte chunks of one row (coalesced reads of the table row and coalesced streaming stores).
template <bool VEC4>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ W, const float* __restrict__ bias, const int32_t* __restrict__ ids,
                   const uint8_t* __restrict__ mask, const float* __restrict__ in_scale, float* __restrict__ xp,
                   int64_t n_tokens, int GH) {
  if (VEC4) {
    const int chunks = GH >> 2;
    const int64_t total = n_tokens * chunks;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t n = i / chunks;
      const int c = (int)(i - n * chunks);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mask[n]) {
        const int32_t id = ids[n];
        v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)id * GH) + c);
        if (in_scale) {
          const float s = in_scale[n];
          v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        }
      }
      if (bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      }
      st_stream_f4(reinterpret_cast<float4*>(xp + n * GH) + c, v);
    }
  } else {
    const int64_t total = n_tokens * GH;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t n = i / GH;
      const int c = (int)(i - n * GH);
      float v = 0.f;
      if (mask[n]) {
        v = __ldg(W + (int64_t)ids[n] * GH + c);
        if (in_scale) v *= in_scale[n];
      }
      if (bias) v += __ldg(bias + c);
      xp[i] = v;
    }
  }
}

extern "C" int seqrec_gather_rows(const float* W_in, const float* b, const int32_t* ids, const uint8_t* mask,
                                  const float* in_scale, float* xp, int64_t n_tokens, int V, int GH, void* stream) {
  SEQREC_ARG(n_tokens > 0 && V > 0 && GH > 0, 1);
  const bool vec = (GH % 4 == 0) && ((reinterpret_cast<uintptr_t>(W_in) | reinterpret_cast<uintptr_t>(xp) |
                                      reinterpret_cast<uintptr_t>(b)) % 16 == 0);
  const int64_t work = vec ? n_tokens * (GH / 4) : n_tokens * GH;
  int blocks = (int)((work + 255) / 256);
  const int cap = SEQREC_NUM_SMS * 16;  // grid-stride: 8 resident CTAs/SM x 2 waves
  if (blocks > cap) blocks = cap;
  if (vec)
    gather_rows_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(W_in, b, ids, mask, in_scale, xp, n_tokens, GH);
  else
    gather_rows_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(W_in, b, ids, mask, in_scale, xp, n_tokens, GH);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// K7: dW[id[n],:] += scale[n]*dxp[n,:].  One warp owns 32 consecutive tokens.  __match_any_sync groups the lanes
// whose tokens hit the same table row; the group's rows are summed in registers by the whole warp and leave as
// ONE vector reduction per 16-byte chunk (warp-aggregated atomics).  The group leader also claims the row in the
// touched[] flag array and appends it to the row list that drives the row-sparse optimizer.
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ dxp, const int32_t* __restrict__ ids,
                        const uint8_t* __restrict__ mask, const float* __restrict__ in_scale,
                        float* __restrict__ dW, int32_t* __restrict__ touched, int32_t* __restrict__ rows,
                        int32_t* __restrict__ n_rows, int64_t n_tokens, int GH) {
  // Work item of a warp = (block of 32 tokens, slab of 128 columns): lane <-> one float4 of the slab (or one float in
  // the scalar variant).  Splitting the row over several warps and batching FOUR id-groups per round keeps 4 x 512 B
  // of independent loads in flight per warp; with one warp walking a whole row group by group the kernel sat at one
  // dependent L2/HBM round trip per 512 B (cfg2: 46 us for 59 MB).
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (GH & 3) == 0;
  const int units = vec ? (GH >> 2) : GH;                    // float4s (or floats) per row
  const int slabs = (units + 31) >> 5;                       // 32-lane slabs per row
  const int64_t n_blocks = (n_tokens + 31) >> 5;
  for (int64_t item = warp_global; item < n_blocks * slabs; item += n_warps) {
    const int64_t base = (item / slabs) * 32;
    const int slab = (int)(item % slabs);
    const int64_t n = base + lane;
    int32_t id = -1;
    float sc = 1.0f;
    if (n < n_tokens && mask[n]) {
      id = ids[n];
      if (in_scale) sc = in_scale[n];
    }
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    const bool leader = (id >= 0) && (lane == (__ffs(peers) - 1));
    if (leader && slab == 0) {                               // one warp per token block claims the rows
      if (atomicExch(touched + id, 1) == 0) {
        const int slot = atomicAdd(n_rows, 1);
        rows[slot] = id;
      }
    }
    unsigned leaders = __ballot_sync(0xffffffffu, leader);
    const int c = slab * 32 + lane;
    const bool on = c < units;
    while (leaders) {
      // up to four groups per round: their first members' loads are independent and issued back to back
      int gl[4];
      unsigned grp[4];
      int32_t gid[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        gl[k] = leaders ? (__ffs(leaders) - 1) : -1;
        if (leaders) leaders &= leaders - 1;
        const int src = gl[k] < 0 ? 0 : gl[k];
        grp[k] = __shfl_sync(0xffffffffu, peers, src);
        gid[k] = __shfl_sync(0xffffffffu, id, src);
      }
      if (vec) {
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gl[k] >= 0) {                                  // first member = the leader itself
            const float s0 = __shfl_sync(0xffffffffu, sc, gl[k]);
            if (on) {
              const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(dxp + (base + gl[k]) * GH) + c);
              acc[k] = make_float4(s0 * v.x, s0 * v.y, s0 * v.z, s0 * v.w);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (gl[k] < 0) continue;                           // warp-uniform
          unsigned g = grp[k] & (grp[k] - 1);                // remaining members (duplicates of the id)
          while (g) {
            const int m = __ffs(g) - 1;
            g &= g - 1;
            const float s1 = __shfl_sync(0xffffffffu, sc, m);
            if (on) {
              const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(dxp + (base + m) * GH) + c);
              acc[k].x += s1 * v.x; acc[k].y += s1 * v.y; acc[k].z += s1 * v.z; acc[k].w += s1 * v.w;
            }
          }
          if (on) red_add_f4(dW + (int64_t)gid[k] * GH + 4 * c, acc[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (gl[k] < 0) continue;
          float acc = 0.f;
          unsigned g = grp[k];
          while (g) {
            const int m = __ffs(g) - 1;
            g &= g - 1;
            const float s1 = __shfl_sync(0xffffffffu, sc, m);
            if (on) acc += s1 * dxp[(base + m) * GH + c];
          }
          if (on) atomicAdd(dW + (int64_t)gid[k] * GH + c, acc);
        }
      }
    }
  }
}

extern "C" int seqrec_scatter_add_rows(const float* dxp, const int32_t* ids, const uint8_t* mask,
                                       const float* in_scale, float* dW_in, int32_t* touched, int32_t* rows,
                                       int32_t* n_rows, int64_t n_tokens, int V, int GH, void* stream) {
  SEQREC_ARG(n_tokens > 0 && V > 0 && GH > 0, 1);
  const int units = (GH & 3) == 0 ? GH / 4 : GH;
  const int64_t warps = ((n_tokens + 31) / 32) * ((units + 31) / 32);   // (token block, column slab) items
  int64_t blocks = (warps + 7) / 8;
  const int cap = SEQREC_NUM_SMS * 16;
  if (blocks > cap) blocks = cap;
  scatter_add_rows_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(dxp, ids, mask, in_scale, dW_in, touched, rows,
                                                                 n_rows, n_tokens, GH);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// Data-parallel helper: after a dense all-reduce of dW_in every rank must update the UNION of touched rows.
__global__ void __launch_bounds__(256)
mark_rows_kernel(const int32_t* __restrict__ ids, const uint8_t* __restrict__ mask, int32_t* __restrict__ touched,
                 int32_t* __restrict__ rows, int32_t* __restrict__ n_rows, int64_t n_tokens) {
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < n_tokens; n += (int64_t)gridDim.x * blockDim.x) {
    if (mask ? (mask[n] != 0) : (ids[n] >= 0)) {
      const int32_t id = ids[n];
      if (id >= 0 && atomicExch(touched + id, 1) == 0) rows[atomicAdd(n_rows, 1)] = id;
    }
  }
}

extern "C" int seqrec_mark_rows(const int32_t* ids, const uint8_t* mask, int32_t* touched, int32_t* rows,
                                int32_t* n_rows, int64_t n_tokens, int V, void* stream) {
  SEQREC_ARG(n_tokens > 0 && V > 0, 1);
  int blocks = (int)((n_tokens + 255) / 256);
  if (blocks > SEQREC_NUM_SMS * 8) blocks = SEQREC_NUM_SMS * 8;
  mark_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(ids, mask, touched, rows, n_rows, n_tokens);
  SEQREC_CHECK_LAUNCH();
  return 0;
}
