// Recurrent weight gradient on the tensor cores:  dU (H, G*H) += sum_n a[n,:]^T . dxp[n,:]  with a = h_{t-1} (or
// r*h_{t-1} for the GRU's candidate block) -- the time-batched GEMM of Theano's scan gradient (model.py:345-352),
// K = all tokens of the batch.
//
// Both operands are token-major in HBM (rows = tokens = the GEMM's K axis), so both enter tcgen05.mma as MN-MAJOR
// operands straight from their bf16 hi/lo images: no transposed copy of the 400 MB dxp tensor is ever made.
// Split-K persistent grid: CTA (tile, split) walks its token range in 64-token stages (TMA, 128-byte swizzle, 3-stage
// ring), accumulates a 128 x 128 fp32 tile in TMEM with the 3-pass hi/lo split (fp32-grade products) and leaves
// through red.global.add.v4.f32 into the pre-zeroed dU.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tma_host.cuh"

namespace {

constexpr int WG_THREADS = 192;        // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int WG_KSTAGE = 64;          // tokens per stage
constexpr uint32_t WG_BLK = 64 * 128;  // one [64 tokens x 64 elements] operand block (8 KB)
constexpr uint32_t WG_STAGE = 8 * WG_BLK;  // A hi (2 blocks) | A lo | D hi | D lo = 64 KB
constexpr int WG_NS = 3;

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                float* __restrict__ C, int ldc, int64_t K, int64_t k_per_cta, int m_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sS = base;                                  // [NS][WG_STAGE]
  const uint32_t sBar = sS + WG_NS * WG_STAGE;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * WG_NS, bar_tfull = sBar + 16 * WG_NS,
                 tmem_slot = bar_tfull + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x % m_tiles, nt = blockIdx.x / m_tiles;
  const int m0 = mt * 128, n0 = nt * 128;
  const int64_t k_begin = (int64_t)blockIdx.y * k_per_cta;
  const int64_t k_end = (k_begin + k_per_cta < K) ? k_begin + k_per_cta : K;
  const int nstages = k_begin < k_end ? (int)((k_end - k_begin + WG_KSTAGE - 1) / WG_KSTAGE) : 0;
  if (nstages == 0) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_NS; ++i) { ptx::mbar_init(bar_full + 8 * i, 1); ptx::mbar_init(bar_empty + 8 * i, 1); }
    ptx::mbar_init(bar_tfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) { ptx::prefetch_tmap(&tmA_hi); ptx::prefetch_tmap(&tmD_hi); }
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < nstages; ++s) {
      ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (ptx::elect_one()) {
        const uint32_t bar = bar_full + 8 * stage;
        const uint32_t dst = sS + stage * WG_STAGE;
        const int k0 = (int)(k_begin + (int64_t)s * WG_KSTAGE);
        ptx::mbar_arrive_expect_tx(bar, WG_STAGE);
        for (int blk = 0; blk < 2; ++blk) {
          ptx::tma_load_2d(dst + (0 + blk) * WG_BLK, &tmA_hi, bar, m0 + blk * 64, k0);
          ptx::tma_load_2d(dst + (2 + blk) * WG_BLK, &tmA_lo, bar, m0 + blk * 64, k0);
          ptx::tma_load_2d(dst + (4 + blk) * WG_BLK, &tmD_hi, bar, n0 + blk * 64, k0);
          ptx::tma_load_2d(dst + (6 + blk) * WG_BLK, &tmD_lo, bar, n0 + blk * 64, k0);
        }
      }
      __syncwarp();
      if (++stage == WG_NS) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // A and B both MN-major (bits 15 and 16): rows of the shared-memory image are K (tokens), 64 M/N elements per row
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < nstages; ++s) {
      ptx::mbar_wait(bar_full + 8 * stage, phase);
      ptx::tc_fence_after_sync();
      const uint32_t st = sS + stage * WG_STAGE;
      const uint64_t a_hi0 = ptx::umma_desc_mn_sw128(st, WG_BLK), a_lo0 = ptx::umma_desc_mn_sw128(st + 2 * WG_BLK, WG_BLK);
      const uint64_t b_hi0 = ptx::umma_desc_mn_sw128(st + 4 * WG_BLK, WG_BLK),
                     b_lo0 = ptx::umma_desc_mn_sw128(st + 6 * WG_BLK, WG_BLK);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < WG_KSTAGE / 16; ++ks) {
          const uint64_t o = (uint64_t)((ks * 16 * 128) >> 4);   // 16 tokens = 16 rows of 128 B further down
          ptx::umma_bf16(tmem_base, a_hi0 + o, b_lo0 + o, idesc, (s > 0 || ks > 0) ? 1u : 0u);
          ptx::umma_bf16(tmem_base, a_lo0 + o, b_hi0 + o, idesc, 1u);
          ptx::umma_bf16(tmem_base, a_hi0 + o, b_hi0 + o, idesc, 1u);
        }
        ptx::umma_commit(bar_empty + 8 * stage);
        if (s == nstages - 1) ptx::umma_commit(bar_tfull);
      }
      __syncwarp();
      if (++stage == WG_NS) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp & 3;                                  // TMEM lane quadrant of this warp
    ptx::mbar_wait(bar_tfull, 0);
    ptx::tc_fence_after_sync();
    float* dst = C + (size_t)(m0 + q * 32 + lane) * ldc + n0;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g)
        red_add_f4(dst + c + 4 * g, make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                                __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])));
    }
    ptx::tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

// C (M rows, ldc) [:, 0:Ncols) += A[0:K, 0:M]^T . D[0:K, 0:Ncols); A / D bf16 hi/lo, row-major with leading
// dimensions lda / ldd (elements)
int launch_wgrad(const uint16_t* A_hi, const uint16_t* A_lo, int lda, const uint16_t* D_hi, const uint16_t* D_lo,
                 int ldd, float* C, int ldc, int M, int Ncols, int64_t K, cudaStream_t st) {
  if (K <= 0) return 0;
  if (M % 128 || Ncols % 128 || (ldc & 3)) return -1040;
  CUtensorMap a_hi, a_lo, d_hi, d_lo;
  int rc;
  if ((rc = tma::make_2d_bf16(&a_hi, A_hi, (uint64_t)K, M, lda, 64, WG_KSTAGE, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tma::make_2d_bf16(&a_lo, A_lo, (uint64_t)K, M, lda, 64, WG_KSTAGE, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tma::make_2d_bf16(&d_hi, D_hi, (uint64_t)K, Ncols, ldd, 64, WG_KSTAGE, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = tma::make_2d_bf16(&d_lo, D_lo, (uint64_t)K, Ncols, ldd, 64, WG_KSTAGE, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  const int m_tiles = M / 128, tiles = m_tiles * (Ncols / 128);
  int splits = (2 * SEQREC_NUM_SMS + tiles - 1) / tiles;     // two waves of CTAs hide the prologue / flush tails
  const int64_t max_splits = (K + 4 * WG_KSTAGE - 1) / (4 * WG_KSTAGE);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  int64_t kpc = (K + splits - 1) / splits;
  kpc = (kpc + WG_KSTAGE - 1) / WG_KSTAGE * WG_KSTAGE;       // stage aligned: only the global tail is ragged
  splits = (int)((K + kpc - 1) / kpc);
  const size_t smem = (size_t)WG_NS * WG_STAGE + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e;
    attr_set = true;
  }
  wgrad_tc_kernel<<<dim3(tiles, splits), WG_THREADS, smem, st>>>(a_hi, a_lo, d_hi, d_lo, C, ldc, K, kpc, m_tiles);
  SEQREC_CHECK_LAUNCH();
  return 0;
}

// out[c] += sum_r in[r, c]   (rows split over blockIdx.y)
__global__ void __launch_bounds__(256)
wg_colsum_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int64_t rows, int cols,
                 int64_t rows_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = blockIdx.y * rows_per_block;
  const int64_t r1 = rows < r0 + rows_per_block ? rows : r0 + rows_per_block;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += in[r * ld + c];
  atomicAdd(out + c, acc);
}

}  // namespace

// dxp_hi/lo: (N, G*H) bf16 split of dxp;  h_hi/lo: (N, H) split of hout;  c_hi/lo: (N, H) split of cst (GRU only).
// dU (H, G*H) and db (G*H) must be pre-zeroed; db is summed from the fp32 dxp.  H must be 128 or 256.
extern "C" int seqrec_rnn_weight_grad_tc(int cell, const float* dxp, const uint16_t* dxp_hi, const uint16_t* dxp_lo,
                                         const uint16_t* h_hi, const uint16_t* h_lo, const uint16_t* c_hi,
                                         const uint16_t* c_lo, float* dU, float* db, int T, int B, int H,
                                         void* stream) {
  SEQREC_ARG(T > 0 && B > 0 && (H == 128 || H == 256), 1);
  SEQREC_ARG(cell == SEQREC_CELL_LSTM || cell == SEQREC_CELL_GRU, 2);
  SEQREC_ARG(dxp && dxp_hi && dxp_lo && h_hi && h_lo && dU && (cell != SEQREC_CELL_GRU || (c_hi && c_lo)), 3);
  cudaStream_t st = as_stream(stream);
  const int G = (cell == SEQREC_CELL_LSTM) ? 4 : 3;
  const int GH = G * H;
  const int64_t N = (int64_t)T * B;
  int rc;
  // blocks fed by h_{t-1}: token n pairs hout[n - B] with dxp[n]
  const int cols = (cell == SEQREC_CELL_GRU) ? 2 * H : GH;
  if ((rc = launch_wgrad(h_hi, h_lo, H, dxp_hi + (size_t)B * GH, dxp_lo + (size_t)B * GH, GH, dU, GH, H, cols, N - B,
                         st)))
    return rc;
  if (cell == SEQREC_CELL_GRU)
    if ((rc = launch_wgrad(c_hi, c_lo, H, dxp_hi + 2 * H, dxp_lo + 2 * H, GH, dU + 2 * H, GH, H, H, N, st))) return rc;
  if (db) {
    int64_t rpb = (N + 63) / 64;
    if (rpb < 64) rpb = 64;
    dim3 grid(ceil_div(GH, 256), ceil_div(N, rpb));
    wg_colsum_kernel<<<grid, 256, 0, st>>>(dxp, GH, db, N, GH, rpb);
    SEQREC_CHECK_LAUNCH();
  }
  return 0;
}
