// Helpers shared by the register-resident scans (rnn_reg.cu: GRU / SimpleRNN, rnn_reg_lstm.cu: LSTM).
#pragma once
#include "common.cuh"

namespace regscan {

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc[g][r] += sum_i vec[r*ldv + k0 + i] * u[g0 + g][i] for NG gates sharing the same vector; the float4 loads of the
// next quad are issued before the FFMAs of the current one (software pipelined, fully unrolled)
template <int RB, int KPT, int NG, int GT>
__device__ __forceinline__ void dot_slices(float (&acc)[NG][RB], const float* __restrict__ vec, int ldv, int k0,
                                           const float (&u)[GT][KPT], int g0) {
  float4 hq[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) hq[r] = *reinterpret_cast<const float4*>(vec + r * ldv + k0);
#pragma unroll
  for (int i = 0; i < KPT; i += 4) {
    float4 hn[RB];
    if (i + 4 < KPT) {
#pragma unroll
      for (int r = 0; r < RB; ++r) hn[r] = *reinterpret_cast<const float4*>(vec + r * ldv + k0 + i + 4);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        acc[g][r] = fmaf(hq[r].x, u[g0 + g][i], acc[g][r]);
        acc[g][r] = fmaf(hq[r].y, u[g0 + g][i + 1], acc[g][r]);
        acc[g][r] = fmaf(hq[r].z, u[g0 + g][i + 2], acc[g][r]);
        acc[g][r] = fmaf(hq[r].w, u[g0 + g][i + 3], acc[g][r]);
      }
    if (i + 4 < KPT) {
#pragma unroll
      for (int r = 0; r < RB; ++r) hq[r] = hn[r];
    }
  }
}

}  // namespace regscan
