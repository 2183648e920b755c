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
__global__ void __launch_bounds__(256)
history_features_kernel(const int32_t* __restrict__ items, const int64_t* __restrict__ offs, float* __restrict__ out,
                        int64_t n_seqs, int T, int V, int freq, const float* __restrict__ table, int table_len,
                        int32_t* __restrict__ err) {
  const int64_t total = n_seqs * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / V;
    const int v = (int)(i - b * V);
    const int64_t start = offs[b], len = offs[b + 1] - start;
    const int64_t rows = len > 0 ? len - 1 : 0;                // xs[:-1]: one row per input position
    const int64_t keep = rows < T ? rows : T;
    const int64_t skip = rows - keep;
    float* o = out + (b * T) * (int64_t)V + v;
    int cnt = 0;
    bool bad = false;
    for (int64_t j = 0; j < skip; ++j) {
      const int32_t it = items[start + j];
      bad |= (it < 0) | (it >= V);
      cnt += (it == v);
    }
    const int pad = T - (int)keep;
    for (int t = 0; t < pad; ++t) o[(int64_t)t * V] = 0.f;
    for (int64_t j = 0; j < keep; ++j) {
      const int32_t it = items[start + skip + j];
      bad |= (it < 0) | (it >= V);
      cnt += (it == v);
      int c = freq ? cnt : (cnt > 0 ? 1 : 0);
      float val = (float)c;
      if (table) val = table[c < table_len ? c : table_len - 1];
      o[(pad + j) * (int64_t)V] = val;
    }
    // (the reference's xi[s] raises IndexError for s >= V and silently wraps a negative s: both are errors here)
    if (v == 0 && bad && err) atomicOr(err, 1);
  }
}

extern "C" int seqrec_history_features(const int32_t* items, const int64_t* offsets, float* c, int64_t n_seqs, int T,
                                       int V, int freq, const float* table, int table_len, int32_t* err,
                                       void* stream) {
  SEQREC_ARG(items && offsets && c && n_seqs > 0 && T > 0 && V > 0, 1);
  SEQREC_ARG(!table || table_len > 1, 2);
  int64_t blocks = (n_seqs * V + 255) / 256;
  if (blocks > SEQREC_NUM_SMS * 8) blocks = SEQREC_NUM_SMS * 8;
  history_features_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(items, offsets, c, n_seqs, T, V, freq, table,
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
