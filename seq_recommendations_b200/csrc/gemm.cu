// Small fp32 SIMT GEMMs around the scan: the dense-feature input projection of RNNBaseline (K2, model.py:245-255),
// the time-batched recurrent weight gradient dU = sum_t hprev_t^T . dxp_t and bias gradient (tail of K4), and a
// transpose for U^T.  These are <2% of the step's flops at every named config; the logits GEMMs that
// dominate live in ce_simt.cu / ce_tc.cu.
#include "common.cuh"

#define GT 64   // tile edge
#define GK 16   // k chunk
#define GP 68   // padded row (multiple of 4 keeps float4 rows aligned)

// C[M,N] (=|+=) A[M,K].B[K,N] (+ bias[N])
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
               const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K, int accumulate) {
  __shared__ __align__(16) float As[GK][GP];
  __shared__ __align__(16) float Bs[GK][GP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += GK) {
    {
      const int r = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + r, k = k0 + kq + i;
        As[kq + i][r] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
      }
      const int kk = tid >> 4, nq = (tid & 15) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + kk, n = n0 + nq + i;
        Bs[kk][nq + i] = (k < K && n < N) ? Bm[(size_t)k * ldb + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (accumulate) v += C[(size_t)m * ldc + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// C[M,N] += A[K,M]^T . B[K,N], K split over blockIdx.z, atomics into a pre-zeroed C
__global__ void __launch_bounds__(256)
gemm_tn_atomic_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                      float* __restrict__ C, int ldc, int M, int N, int K, int k_per_split) {
  __shared__ __align__(16) float As[GK][GP];
  __shared__ __align__(16) float Bs[GK][GP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int kb = blockIdx.z * k_per_split;
  const int ke = min(K, kb + k_per_split);
  float acc[4][4] = {};
  for (int k0 = kb; k0 < ke; k0 += GK) {
    const int kk = tid >> 4, q = (tid & 15) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + kk;
      As[kk][q + i] = (k < ke && m0 + q + i < M) ? A[(size_t)k * lda + m0 + q + i] : 0.f;
      Bs[kk][q + i] = (k < ke && n0 + q + i < N) ? Bm[(size_t)k * ldb + n0 + q + i] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(C + (size_t)m * ldc + n, acc[i][j]);
    }
  }
}

// out[c] += sum_r in[r, c]   (rows split over blockIdx.y)
__global__ void __launch_bounds__(256)
colsum_atomic_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int64_t rows, int cols,
                     int64_t rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += in[r * ld + c];
  atomicAdd(out + c, acc);
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

static int launch_gemm_tn(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, int K,
                          cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int tiles = ceil_div(M, GT) * ceil_div(N, GT);
  int splits = ceil_div(2 * SEQREC_NUM_SMS, tiles);
  const int max_splits = ceil_div(K, 4 * GK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int kps = ceil_div(K, splits);
  kps = ceil_div(kps, GK) * GK;
  splits = ceil_div(K, kps);
  dim3 grid(ceil_div(N, GT), ceil_div(M, GT), splits);
  gemm_tn_atomic_kernel<<<grid, 256, 0, st>>>(A, lda, Bm, ldb, C, ldc, M, N, K, kps);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

extern "C" int seqrec_gemm_nn(const float* A, const float* Bm, const float* bias, float* C, int M, int N, int K,
                              int accumulate, void* stream) {
  SEQREC_ARG(M > 0 && N > 0 && K > 0, 1);
  dim3 grid(ceil_div(N, GT), ceil_div(M, GT));
  gemm_nn_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, K, Bm, N, bias, C, N, M, N, K, accumulate);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

extern "C" int seqrec_gemm_tn_atomic(const float* A, const float* Bm, float* C, int M, int N, int K, void* stream) {
  SEQREC_ARG(M > 0 && N > 0 && K > 0, 1);
  return launch_gemm_tn(A, M, Bm, N, C, N, M, N, K, as_stream(stream));
}

extern "C" int seqrec_transpose(const float* in, float* out, int rows, int cols, void* stream) {
  SEQREC_ARG(rows > 0 && cols > 0, 1);
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(in, out, rows, cols);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// dU (H, G*H) += sum_{t>=1} hout[t-1]^T . dxp[t]  (h_{-1} = 0 contributes nothing);  GRU candidate block uses
// cst = r*h_{t-1} at the same t.  db += column sums of dxp.
extern "C" int seqrec_rnn_weight_grad(int cell, const float* dxp, const float* hout, const float* cst, float* dU,
                                      float* db, int T, int B, int H, void* stream) {
  SEQREC_ARG(T > 0 && B > 0 && H > 0, 1);
  cudaStream_t st = as_stream(stream);
  const int G = (cell == SEQREC_CELL_LSTM) ? 4 : (cell == SEQREC_CELL_GRU ? 3 : 1);
  const int GH = G * H;
  const int64_t N = (int64_t)T * B;
  int rc = 0;
  if (T > 1) {
    const int Kt = (int)(N - B);
    const int cols = (cell == SEQREC_CELL_GRU) ? 2 * H : GH;
    rc = launch_gemm_tn(hout, H, dxp + (size_t)B * GH, GH, dU, GH, H, cols, Kt, st);
    if (rc) return rc;
  }
  if (cell == SEQREC_CELL_GRU) {
    rc = launch_gemm_tn(cst, H, dxp + 2 * H, GH, dU + 2 * H, GH, H, H, (int)N, st);
    if (rc) return rc;
  }
  if (db) {
    int64_t rpb = (N + 63) / 64;
    if (rpb < 64) rpb = 64;
    dim3 grid(ceil_div(GH, 256), ceil_div(N, rpb));
    colsum_atomic_kernel<<<grid, 256, 0, st>>>(dxp, GH, db, N, GH, rpb);
    SEQREC_CHECK_LAUNCH();
  }
  return 0;
}


// out[t,b,:] = in[t,b,:] * m[b,:]   (h_{t-1} * rm[g]: operand of gate block g's dU under recurrent dropout)
__global__ void __launch_bounds__(256)
mul_rows_bcast_kernel(const float* __restrict__ in, const float* __restrict__ m, float* __restrict__ out, int64_t total,
                      int64_t BH) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) out[i] = in[i] * m[i % BH];
}

// seqrec_rnn_weight_grad under recurrent dropout: gate block g sees h_{t-1} * rec_mask[g], so dU is G separate products
// (GRU candidate block: cst = r * h_{t-1} * rec_mask[2], written by seqrec_rnn_backward_rd).  scratch: T*B*H floats.
extern "C" int seqrec_rnn_weight_grad_rd(int cell, const float* dxp, const float* hout, const float* cst,
                                         const float* rec_mask, float* scratch, float* dU, float* db, int T, int B,
                                         int H, void* stream) {
  SEQREC_ARG(T > 0 && B > 0 && H > 0 && rec_mask && scratch, 1);
  cudaStream_t st = as_stream(stream);
  const int G = (cell == SEQREC_CELL_LSTM) ? 4 : (cell == SEQREC_CELL_GRU ? 3 : 1);
  const int GH = G * H;
  const int64_t N = (int64_t)T * B;
  int rc = 0;
  const int direct = (cell == SEQREC_CELL_GRU) ? 2 : G;      // gate blocks whose operand is h_{t-1} * rm[g]
  if (T > 1) {
    const int64_t total = (N - B) * H;
    int64_t blocks = (total + 255) / 256;
    if (blocks > SEQREC_NUM_SMS * 16) blocks = SEQREC_NUM_SMS * 16;
    for (int g = 0; g < direct; ++g) {
      mul_rows_bcast_kernel<<<(int)blocks, 256, 0, st>>>(hout, rec_mask + (size_t)g * B * H, scratch, total,
                                                         (int64_t)B * H);
      SEQREC_CHECK_LAUNCH();
      rc = launch_gemm_tn(scratch, H, dxp + (size_t)B * GH + g * H, GH, dU + g * H, GH, H, H, (int)(N - B), st);
      if (rc) return rc;
    }
  }
  if (cell == SEQREC_CELL_GRU) {
    rc = launch_gemm_tn(cst, H, dxp + 2 * H, GH, dU + 2 * H, GH, H, H, (int)N, st);
    if (rc) return rc;
  }
  if (db) {
    int64_t rpb = (N + 63) / 64;
    if (rpb < 64) rpb = 64;
    dim3 grid(ceil_div(GH, 256), ceil_div(N, rpb));
    colsum_atomic_kernel<<<grid, 256, 0, st>>>(dxp, GH, db, N, GH, rpb);
    SEQREC_CHECK_LAUNCH();
  }
  return 0;
}
