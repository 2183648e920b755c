// K2 on the 5th-generation tensor cores: the time-batched dense input projection  xp (+)= X . W_in  (RNNBaseline with
// [onehot || xs], model.py:245-255; the x_to_z branch of RNNFullModel, model.py:354-358) -- and, with the operands staged
// accordingly, every other plain product of the history-feature branches (x . B, dZ . W^T, X^T . dZ: model.py:376-392).
//
//   C[M,N] (+)= A[M,K] . Bt[N,K]^T (+ bias[N])        fp32 in / out, fp32 accumulate
//
// Operands are bf16 hi/lo pairs staged K-major by seqrec_split_bf16 (transpose as needed); "fp32 mode" issues the
// 3-pass split product a_hi.b_lo + a_lo.b_hi + a_hi.b_hi (~2^-16 relative), like the logits kernels of ce_tc.cu.
//
// sm_100a design: persistent grid (one CTA per SM) over 128 x 128 output tiles; warp 0 = TMA producer
// (cp.async.bulk.tensor, 128-byte swizzle, 64-element K blocks of BOTH operands through one ring of NS stages), warp 1 =
// MMA issuer (tcgen05.mma, M = 128, accumulators double-buffered in TMEM so the epilogue of tile i overlaps the
// mainloop of tile i+1), warps 2..5 = epilogue (tcgen05.ld 32x32b: one thread owns one output row; bias / accumulate
// applied in registers, 16-byte stores).  Ragged M / N / K edges are zero-filled by the TMA unit.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_host.cuh"

namespace {

constexpr int GM = 128, GN = 128, GKB = 64;
constexpr int G_TILE_B = 128 * 128;       // bytes of one [128 rows x 64 bf16] operand block
constexpr int G_THREADS = 192;

template <bool X3>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const float* __restrict__ bias, float* __restrict__ C, int64_t M, int N, int K, int64_t ldc,
               int accumulate) {
  constexpr int NP = X3 ? 2 : 1;
  constexpr int NS = X3 ? 3 : 6;
  constexpr uint32_t STAGE_B = 2 * NP * G_TILE_B;             // A hi [lo] | B hi [lo]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sRing = base;
  const uint32_t sBar = sRing + NS * STAGE_B;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * NS, bar_tfull = sBar + 16 * NS, bar_tempty = bar_tfull + 16,
                 tmem_slot = bar_tempty + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_mt = (int)((M + GM - 1) / GM), n_nt = (N + GN - 1) / GN, n_kb = (K + GKB - 1) / GKB;
  const int64_t n_tiles = (int64_t)n_mt * n_nt;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { ptx::mbar_init(bar_full + 8 * i, 1); ptx::mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_tfull + 8 * i, 1); ptx::mbar_init(bar_tempty + 8 * i, 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) { ptx::prefetch_tmap(&tmA_hi); ptx::prefetch_tmap(&tmB_hi); }
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (int)(t / n_nt) * GM, n0 = (int)(t % n_nt) * GN;
      for (int kb = 0; kb < n_kb; ++kb) {
        ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (ptx::elect_one()) {
          const uint32_t fb = bar_full + 8 * stage;
          ptx::mbar_arrive_expect_tx(fb, STAGE_B);
          const uint32_t dst = sRing + stage * STAGE_B;
          ptx::tma_load_2d(dst, &tmA_hi, fb, kb * GKB, m0);
          if (X3) ptx::tma_load_2d(dst + G_TILE_B, &tmA_lo, fb, kb * GKB, m0);
          ptx::tma_load_2d(dst + NP * G_TILE_B, &tmB_hi, fb, kb * GKB, n0);
          if (X3) ptx::tma_load_2d(dst + (NP + 1) * G_TILE_B, &tmB_lo, fb, kb * GKB, n0);
        }
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(GM, GN);
    int stage = 0, tc = 0;
    uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tc) {
      const int buf = tc & 1;
      ptx::mbar_wait(bar_tempty + 8 * buf, ((tc >> 1) & 1) ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d = tmem_base + buf * GN;
      for (int kb = 0; kb < n_kb; ++kb) {
        ptx::mbar_wait(bar_full + 8 * stage, phase);
        ptx::tc_fence_after_sync();
        const uint32_t a = sRing + stage * STAGE_B, b = a + NP * G_TILE_B;
        const uint64_t da_hi = ptx::umma_desc_k_sw128(a), da_lo = ptx::umma_desc_k_sw128(a + G_TILE_B);
        const uint64_t db_hi = ptx::umma_desc_k_sw128(b), db_lo = ptx::umma_desc_k_sw128(b + G_TILE_B);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < GKB / 16; ++k) {
            const int e = k * 16;
            const uint32_t acc = (kb == 0 && k == 0) ? 0u : 1u;
            if (X3) {
              ptx::umma_bf16(d, ptx::umma_desc_advance_k(da_hi, e), ptx::umma_desc_advance_k(db_lo, e), idesc, acc);
              ptx::umma_bf16(d, ptx::umma_desc_advance_k(da_lo, e), ptx::umma_desc_advance_k(db_hi, e), idesc, 1u);
              ptx::umma_bf16(d, ptx::umma_desc_advance_k(da_hi, e), ptx::umma_desc_advance_k(db_hi, e), idesc, 1u);
            } else {
              ptx::umma_bf16(d, ptx::umma_desc_advance_k(da_hi, e), ptx::umma_desc_advance_k(db_hi, e), idesc, acc);
            }
          }
          ptx::umma_commit(bar_empty + 8 * stage);            // this stage may be refilled once the MMAs have read it
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
      if (ptx::elect_one()) ptx::umma_commit(bar_tfull + 8 * buf);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------------------------------- epilogue
    const int q = warp & 3;                                    // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
    int tc = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tc) {
      const int buf = tc & 1;
      const int64_t m = (t / n_nt) * GM + row;
      const int n0 = (int)(t % n_nt) * GN;
      ptx::mbar_wait(bar_tfull + 8 * buf, (tc >> 1) & 1);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < GN; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * GN + c, r);
        ptx::tmem_ld_wait();
        if (m >= M || n0 + c >= N) continue;
        float* dst = C + m * ldc + n0 + c;
        const int valid = min(32, N - (n0 + c));
        if (vec_ok && valid == 32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 v = make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                   __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
            if (bias) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + c) + g);
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            }
            if (accumulate) {
              const float4 o = *(reinterpret_cast<const float4*>(dst) + g);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *(reinterpret_cast<float4*>(dst) + g) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < valid) {
              float v = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + c + j) : 0.f);
              if (accumulate) v += dst[j];
              dst[j] = v;
            }
          }
        }
      }
      ptx::tc_fence_before_sync();                             // accumulator read: hand the buffer back
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8 * buf);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

template <bool X3>
int launch_gemm_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
                   const float* bias, float* C, int64_t M, int N, int K, int64_t ldc, int accumulate,
                   cudaStream_t st) {
  constexpr int NP = X3 ? 2 : 1;
  constexpr int NS = X3 ? 3 : 6;
  const size_t smem = 1024 + (size_t)NS * 2 * NP * G_TILE_B + 256;
  auto k = gemm_tc_kernel<X3>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -(int)e;
  const int64_t tiles = ((M + GM - 1) / GM) * ((N + GN - 1) / GN);
  const int grid = (int)(tiles < SEQREC_NUM_SMS ? tiles : SEQREC_NUM_SMS);
  k<<<grid, G_THREADS, smem, st>>>(a_hi, a_lo, b_hi, b_lo, bias, C, M, N, K, ldc, accumulate);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int seqrec_gemm_tc(const uint16_t* A_hi, const uint16_t* A_lo, const uint16_t* Bt_hi, const uint16_t* Bt_lo,
                              const float* bias, float* C, int64_t M, int N, int K, int64_t lda, int64_t ldb,
                              int64_t ldc, int accumulate, int x3, void* stream) {
  SEQREC_ARG(M > 0 && N > 0 && K > 0 && lda >= K && ldb >= K && ldc >= N, 1);
  SEQREC_ARG(A_hi && Bt_hi && C && (!x3 || (A_lo && Bt_lo)), 2);
  SEQREC_ARG(lda % 8 == 0 && ldb % 8 == 0, 3);               // TMA: 16-byte row pitch
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  int rc;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if ((rc = tma::make_2d_bf16(&a_hi, A_hi, M, K, lda, GKB, GM, sw))) return rc;
  if ((rc = tma::make_2d_bf16(&a_lo, x3 ? A_lo : A_hi, M, K, lda, GKB, GM, sw))) return rc;
  if ((rc = tma::make_2d_bf16(&b_hi, Bt_hi, N, K, ldb, GKB, GN, sw))) return rc;
  if ((rc = tma::make_2d_bf16(&b_lo, x3 ? Bt_lo : Bt_hi, N, K, ldb, GKB, GN, sw))) return rc;
  cudaStream_t st = as_stream(stream);
  return x3 ? launch_gemm_tc<true>(a_hi, a_lo, b_hi, b_lo, bias, C, M, N, K, ldc, accumulate, st)
            : launch_gemm_tc<false>(a_hi, a_lo, b_hi, b_lo, bias, C, M, N, K, ldc, accumulate, st);
}
