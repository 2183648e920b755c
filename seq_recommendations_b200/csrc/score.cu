// K9: top-k next-item scoring without materialising the (N,V) probabilities.
//
// The reference only has `model.predict` -> full (N,T,V) softmax (model.py:194-195) consumed by p(true item)
// (model.py:106-112); there is no ranking code.  Top-k is defined here as the k items with the largest
// logits per row, ties broken by the LOWER item id -- the order a stable argsort of -p gives.  Softmax is monotone,
// so ranking logits equals ranking probabilities; probabilities are reported as exp(z-m)/s from the K5 statistics.
//
// Tiling as in ce_simt.cu (64 rows x 64 items, fp32 FFMA, exact logits).  Each warp owns 8 rows of the tile and keeps
// their running top-k lists sorted in shared memory; a candidate is inserted only if it beats the current k-th.
#include "common.cuh"

#define ST 64
#define SP 68
#define SK_MAX 64
#define S_THREADS 256

__global__ void __launch_bounds__(S_THREADS)
topk_simt_kernel(const float* __restrict__ hout, const float* __restrict__ W_out, const float* __restrict__ b_out,
                 const float* __restrict__ mrow, const float* __restrict__ srow, int32_t* __restrict__ topk_ids,
                 float* __restrict__ topk_p, int64_t n_rows, int H, int V, int k) {
  extern __shared__ __align__(16) float smem[];
  float* A_s = smem;                         // [H][SP]
  float* W_s = A_s + (size_t)H * SP;         // [H][SP]
  float* Z_s = W_s + (size_t)H * SP;         // [ST][SP]
  float* tv = Z_s + ST * SP;                 // [ST][SK_MAX] values, sorted descending
  int32_t* ti = reinterpret_cast<int32_t*>(tv + ST * SK_MAX);  // [ST][SK_MAX] ids
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * ST;

  for (int i = tid; i < ST * H; i += S_THREADS) {
    const int r = i / H, h = i - r * H;
    A_s[h * SP + r] = (n0 + r < n_rows) ? hout[(n0 + r) * H + h] : 0.f;
  }
  for (int i = tid; i < ST * SK_MAX; i += S_THREADS) { tv[i] = -INFINITY; ti[i] = 0x7fffffff; }

  for (int v0 = 0; v0 < V; v0 += ST) {
    __syncthreads();
    for (int i = tid; i < ST * H; i += S_THREADS) {
      const int h = i / ST, c = i - h * ST;
      W_s[h * SP + c] = (v0 + c < V) ? W_out[(size_t)h * V + v0 + c] : 0.f;
    }
    __syncthreads();
    float acc[4][4] = {};
#pragma unroll 4
    for (int h = 0; h < H; ++h) {
      const float4 a = *reinterpret_cast<const float4*>(A_s + h * SP + ty * 4);
      const float4 b = *reinterpret_cast<const float4*>(W_s + h * SP + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int v = v0 + tx * 4 + j;
      const float bj = (b_out && v < V) ? b_out[v] : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) Z_s[(ty * 4 + i) * SP + tx * 4 + j] = (v < V) ? acc[i][j] + bj : -INFINITY;
    }
    __syncthreads();
    // each warp merges the tile into the lists of its 8 rows; candidates are visited in ascending item id
    for (int rr = 0; rr < 8; ++rr) {
      const int row = warp * 8 + rr;
      float* rv = tv + row * SK_MAX;
      int32_t* ri = ti + row * SK_MAX;
      for (int half = 0; half < 2; ++half) {
        const float zc = Z_s[row * SP + half * 32 + lane];
        float thr = rv[k - 1];
        unsigned cand = __ballot_sync(0xffffffffu, zc > thr);
        while (cand) {
          const int l = __ffs(cand) - 1;
          cand &= cand - 1;
          const float z = __shfl_sync(0xffffffffu, zc, l);
          if (!(z > thr)) continue;  // the threshold may have risen since the ballot (uniform branch)
          const int32_t id = v0 + half * 32 + l;
          // insert position = number of entries >= z (keeps earlier, lower-id equals ahead)
          const float e0 = rv[lane];
          const float e1 = (lane + 32 < SK_MAX) ? rv[lane + 32] : -INFINITY;
          const int32_t i0 = ri[lane], i1 = ri[lane + 32];
          const int pos = __popc(__ballot_sync(0xffffffffu, e0 >= z)) + __popc(__ballot_sync(0xffffffffu, e1 >= z));
          __syncwarp();
          if (lane >= pos && lane + 1 < k) { rv[lane + 1] = e0; ri[lane + 1] = i0; }
          if (lane + 32 >= pos && lane + 33 < k) { rv[lane + 33] = e1; ri[lane + 33] = i1; }
          if (lane == 0) { rv[pos] = z; ri[pos] = id; }
          __syncwarp();
          thr = rv[k - 1];
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < ST * k; i += S_THREADS) {
    const int r = i / k, j = i - r * k;
    const int64_t n = n0 + r;
    if (n >= n_rows) continue;
    const float z = tv[r * SK_MAX + j];
    topk_ids[n * k + j] = ti[r * SK_MAX + j];
    if (topk_p) topk_p[n * k + j] = (mrow && srow) ? expf(z - mrow[n]) / srow[n] : z;
  }
}

extern "C" int seqrec_topk(const float* hout, const float* W_out, const float* b_out, const float* m, const float* s,
                           int32_t* topk_ids, float* topk_p, int64_t n_rows, int H, int V, int k, void* stream) {
  SEQREC_ARG(n_rows > 0 && H > 0 && V > 0, 1);
  SEQREC_ARG(k >= 1 && k <= SK_MAX && k <= V, 2);
  const size_t smem = sizeof(float) * ((size_t)2 * H * SP + ST * SP + 2 * ST * SK_MAX);
  if (smem > 227 * 1024) return -1010;
  cudaError_t e = cudaFuncSetAttribute(topk_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  topk_simt_kernel<<<ceil_div(n_rows, ST), S_THREADS, smem, as_stream(stream)>>>(hout, W_out, b_out, m, s, topk_ids,
                                                                                 topk_p, n_rows, H, V, k);
  SEQREC_CHECK_LAUNCH();
  return 0;
}
