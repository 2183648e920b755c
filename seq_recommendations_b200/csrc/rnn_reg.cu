// K3 / K4 fast path: GRU and SimpleRNN scans with the recurrent kernel held in REGISTERS for all T timesteps.
//
// For H <= 128 the whole recurrent kernel U (H x G*H fp32; cfg2 GRU-128: 192 KB) fits in the 256 KB register file of
// one SM.  A persistent CTA of 4*CG threads (CG = H rounded up to 32) owns RB batch rows for the whole sequence:
// thread (s, c) keeps the K-slice s of the gate columns of hidden unit c (forward) or of row c of U (backward) in
// registers, so a timestep reads only the RB hidden vectors from shared memory (warp-wide broadcasts), does its share
// of the (RB x H).(H x G*H) product as register FFMAs, and the four K-slices are summed through shared memory.
// Same semantics, inputs and outputs as the generic kernels in rnn_scan.cu (Keras-2.0.x GRU: reset applied BEFORE the
// recurrent matmul -> two dependent matvec phases per step; Theano K.rnn mask switch).
#include "common.cuh"

namespace {

constexpr int KS = 4;  // K-slices

template <int CELL>
struct Gates { static constexpr int G = (CELL == SEQREC_CELL_GRU) ? 3 : 1; };

// acc[r] += sum_i vec[r*ldv + k0 + i] * u[i]   (vec in shared memory, read as warp-wide float4 broadcasts)
template <int RB, int KPT>
__device__ __forceinline__ void dot_slice(float (&acc)[RB], const float* __restrict__ vec, int ldv, int k0,
                                          const float (&u)[KPT]) {
#pragma unroll
  for (int i = 0; i < KPT; i += 4) {
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float4 h = *reinterpret_cast<const float4*>(vec + r * ldv + k0 + i);
      acc[r] = fmaf(h.x, u[i], acc[r]);
      acc[r] = fmaf(h.y, u[i + 1], acc[r]);
      acc[r] = fmaf(h.z, u[i + 2], acc[r]);
      acc[r] = fmaf(h.w, u[i + 3], acc[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int CELL, int ACT, int RB, int KPT>
__global__ void __launch_bounds__(KS * 128, 1)
rnn_forward_reg_kernel(float* __restrict__ xg, const float* __restrict__ U, const uint8_t* __restrict__ mask,
                       float* __restrict__ hout, int T, int B, int H) {
  constexpr int G = Gates<CELL>::G;
  constexpr int KP = KS * KPT;                 // padded hidden size seen by the dot products
  const int CG = blockDim.x / KS;
  const int GH = G * H;
  extern __shared__ __align__(16) float smem[];
  float* h_s = smem;                           // [RB][KP]  h_{t-1}
  float* rh_s = h_s + RB * KP;                 // [RB][KP]  GRU: r * h_{t-1}
  float* part_s = rh_s + RB * KP;              // [KS][RB][G][CG] partial sums
  const int tid = threadIdx.x;
  const int s = tid / CG, c = tid - s * CG;
  const int b0 = blockIdx.x * RB;
  const bool owner = (s == 0) && (c < H);      // does the gate math of hidden unit c

  float u[G][KPT];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const int k = s * KPT + i;
      u[g][i] = (k < H && c < H) ? U[(size_t)k * GH + g * H + c] : 0.f;
    }
  for (int i = tid; i < 2 * RB * KP; i += blockDim.x) h_s[i] = 0.f;   // h_s and rh_s
  // Everything a step reads from global memory is prefetched one step ahead, so no load latency sits on the
  // sequential critical path.
  float xpre[RB][G];                           // input projection of the current step (owners only)
  bool mpre[RB];                               // mask of the current step
  float hreg[RB];                              // owner's copy of h_{t-1}[c]
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    hreg[r] = 0.f;
    mpre[r] = (owner && b0 + r < B) ? (mask[(size_t)b0 + r] != 0) : false;
#pragma unroll
    for (int g = 0; g < G; ++g)
      xpre[r][g] = (owner && b0 + r < B) ? xg[((size_t)b0 + r) * GH + g * H + c] : 0.f;
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const size_t tok0 = (size_t)t * B + b0;
    constexpr int G1 = (CELL == SEQREC_CELL_GRU) ? 2 : 1;   // gates whose recurrent product uses h_{t-1} itself
    // ---- phase 1: partial h.U over this thread's K-slice
#pragma unroll
    for (int g = 0; g < G1; ++g) {
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = 0.f;
      dot_slice<RB, KPT>(acc, h_s, KP, s * KPT, u[g]);
#pragma unroll
      for (int r = 0; r < RB; ++r) part_s[((s * RB + r) * G + g) * CG + c] = acc[r];
    }
    __syncthreads();
    float zreg[RB], rreg[RB];
    if (CELL == SEQREC_CELL_GRU) {
      if (owner) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float az = xpre[r][0], ar = xpre[r][1];
#pragma unroll
          for (int q = 0; q < KS; ++q) {
            az += part_s[((q * RB + r) * G + 0) * CG + c];
            ar += part_s[((q * RB + r) * G + 1) * CG + c];
          }
          zreg[r] = hard_sigmoid_f(az);
          rreg[r] = hard_sigmoid_f(ar);
          rh_s[r * KP + c] = rreg[r] * hreg[r];
        }
      }
      __syncthreads();
      // ---- phase 2: candidate, (r*h).U_h
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = 0.f;
      dot_slice<RB, KPT>(acc, rh_s, KP, s * KPT, u[2]);
#pragma unroll
      for (int r = 0; r < RB; ++r) part_s[((s * RB + r) * G + 2) * CG + c] = acc[r];
      __syncthreads();
    }
    // ---- phase 3: gate math, state update under the mask, stores; prefetch the next step's input projection
    if (owner) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (b0 + r >= B) continue;
        const size_t tok = tok0 + r;
        const bool m = mpre[r];
        float hn;
        float* gp = xg + tok * GH;
        if (CELL == SEQREC_CELL_GRU) {
          float ah = xpre[r][2];
#pragma unroll
          for (int q = 0; q < KS; ++q) ah += part_s[((q * RB + r) * G + 2) * CG + c];
          const float hh = act_f<ACT>(ah);
          hn = zreg[r] * hreg[r] + (1.0f - zreg[r]) * hh;
          gp[c] = zreg[r]; gp[H + c] = rreg[r]; gp[2 * H + c] = hh;
        } else {
          float a = xpre[r][0];
#pragma unroll
          for (int q = 0; q < KS; ++q) a += part_s[((q * RB + r) * G + 0) * CG + c];
          hn = act_f<ACT>(a);
          gp[c] = hn;
        }
        const float hv = m ? hn : hreg[r];
        hreg[r] = hv;
        h_s[r * KP + c] = hv;
        hout[tok * H + c] = hv;
        if (t + 1 < T) {
          mpre[r] = mask[tok + B] != 0;
#pragma unroll
          for (int g = 0; g < G; ++g) xpre[r][g] = xg[(tok + B) * GH + g * H + c];
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward: thread (s, c) holds U[c][g*H + s*KPT + i] -- row c of U, K-slice s of every gate block.
template <int CELL, int ACT, int RB, int KPT>
__global__ void __launch_bounds__(KS * 128, 1)
rnn_backward_reg_kernel(float* __restrict__ xg, const float* __restrict__ U, const uint8_t* __restrict__ mask,
                        const float* __restrict__ hout, float* __restrict__ cst, const float* __restrict__ dhout,
                        int T, int B, int H) {
  constexpr int G = Gates<CELL>::G;
  constexpr int KP = KS * KPT;
  const int CG = blockDim.x / KS;
  const int GH = G * H;
  extern __shared__ __align__(16) float smem[];
  float* da_s = smem;                          // [G][RB][KP] pre-activation gradients of this step
  float* part_s = da_s + G * RB * KP;          // [KS][RB][CG]
  const int tid = threadIdx.x;
  const int s = tid / CG, c = tid - s * CG;
  const int b0 = blockIdx.x * RB;
  const bool owner = (s == 0) && (c < H);

  float u[G][KPT];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
      const int j = s * KPT + i;
      u[g][i] = (j < H && c < H) ? U[(size_t)c * GH + g * H + j] : 0.f;
    }
  for (int i = tid; i < G * RB * KP; i += blockDim.x) da_s[i] = 0.f;
  float dh_carry[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) dh_carry[r] = 0.f;
  // Per-step global inputs of the owners, prefetched one step ahead (no load latency on the sequential path):
  // dL/dhout, the saved gates (GRU: z, r, hh; SimpleRNN: the output y), h_{t-1} and the mask.
  struct StepIn { float dh, g0, g1, g2, hprev; bool m; };
  auto load_step = [&](int t, int r) {
    StepIn in;
    in.dh = 0.f; in.g0 = 0.f; in.g1 = 0.f; in.g2 = 0.f; in.hprev = 0.f; in.m = false;
    if (owner && b0 + r < B && t >= 0) {
      const size_t tok = (size_t)t * B + b0 + r;
      in.m = mask[tok] != 0;
      in.dh = dhout[tok * H + c];
      if (CELL == SEQREC_CELL_GRU) {
        const float* gp = xg + tok * GH;
        in.g0 = gp[c]; in.g1 = gp[H + c]; in.g2 = gp[2 * H + c];
        in.hprev = (t > 0) ? hout[(tok - B) * H + c] : 0.f;
      } else {
        in.g0 = hout[tok * H + c];
      }
    }
    return in;
  };
  StepIn pre[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) pre[r] = load_step(T - 1, r);
  __syncthreads();

  for (int t = T - 1; t >= 0; --t) {
    const size_t tok0 = (size_t)t * B + b0;
    float direct[RB], hp[RB], rr[RB];
    bool on[RB];
    // ---- phase 1 (owners): elementwise gate gradients
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      direct[r] = 0.f; hp[r] = 0.f; rr[r] = 0.f; on[r] = false;
      const StepIn in = pre[r];
      pre[r] = load_step(t - 1, r);            // in flight behind this step's matvec phases
      if (!owner || b0 + r >= B) continue;
      const size_t tok = tok0 + r;
      const bool m = in.m;
      const float dh = in.dh + dh_carry[r];
      on[r] = m;
      if (!m) {
        direct[r] = dh;                        // masked step: h_t = h_{t-1}, no gate gradient
#pragma unroll
        for (int g = 0; g < G; ++g) da_s[(g * RB + r) * KP + c] = 0.f;
        if (CELL == SEQREC_CELL_GRU) cst[tok * H + c] = 0.f;
        continue;
      }
      if (CELL == SEQREC_CELL_GRU) {
        const float z = in.g0, rg = in.g1, hh = in.g2;
        const float hprev = in.hprev;
        da_s[(0 * RB + r) * KP + c] = dh * (hprev - hh) * hard_sigmoid_grad_from_y(z);
        da_s[(2 * RB + r) * KP + c] = dh * (1.0f - z) * act_grad_from_y<ACT>(hh);
        direct[r] = dh * z;
        hp[r] = hprev;
        rr[r] = rg;
        cst[tok * H + c] = rg * hprev;         // operand of dU's candidate block
      } else {
        da_s[(0 * RB + r) * KP + c] = dh * act_grad_from_y<ACT>(in.g0);
      }
    }
    __syncthreads();
    if (CELL == SEQREC_CELL_GRU) {
      // ---- phase 2: d(r*h_{t-1})[c] = sum_j da_h[j] * U[c][2H + j]
      {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.f;
        dot_slice<RB, KPT>(acc, da_s + 2 * RB * KP, KP, s * KPT, u[2]);
#pragma unroll
        for (int r = 0; r < RB; ++r) part_s[(s * RB + r) * CG + c] = acc[r];
      }
      __syncthreads();
      if (owner) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float d_rh = 0.f;
#pragma unroll
          for (int q = 0; q < KS; ++q) d_rh += part_s[(q * RB + r) * CG + c];
          if (on[r]) {
            da_s[(1 * RB + r) * KP + c] = d_rh * hp[r] * hard_sigmoid_grad_from_y(rr[r]);
            direct[r] += d_rh * rr[r];
          }
        }
      }
      __syncthreads();
    }
    // ---- phase 3: dh_{t-1}[c] = direct + sum_j da[j] * U[c][j] over the gates that see h_{t-1} directly
    {
      constexpr int G1 = (CELL == SEQREC_CELL_GRU) ? 2 : 1;
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = 0.f;
#pragma unroll
      for (int g = 0; g < G1; ++g) dot_slice<RB, KPT>(acc, da_s + g * RB * KP, KP, s * KPT, u[g]);
      // phase 2's partials were consumed before the barrier above, so part_s can be reused
#pragma unroll
      for (int r = 0; r < RB; ++r) part_s[(s * RB + r) * CG + c] = acc[r];
    }
    __syncthreads();
    if (owner) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (b0 + r >= B) continue;
        float v = direct[r];
#pragma unroll
        for (int q = 0; q < KS; ++q) v += part_s[(q * RB + r) * CG + c];
        dh_carry[r] = v;
        // dxp[t] = da (overwrites the saved gates, already consumed in phase 1)
        float* gp = xg + (tok0 + r) * GH;
#pragma unroll
        for (int g = 0; g < G; ++g) gp[g * H + c] = da_s[(g * RB + r) * KP + c];
      }
    }
    __syncthreads();
  }
}

template <int CELL, int ACT, int RB, int KPT>
int launch_pair(bool fwd, float* xg, const float* U, const uint8_t* mask, float* hout, float* cst, const float* dhout,
                int T, int B, int H, cudaStream_t st) {
  constexpr int G = Gates<CELL>::G;
  const int CG = (H + 31) / 32 * 32;
  const int threads = KS * CG;
  const int grid = ceil_div(B, RB);
  constexpr int KP = KS * KPT;
  if (fwd) {
    const size_t smem = sizeof(float) * (size_t)(2 * RB * KP + KS * RB * G * CG);
    rnn_forward_reg_kernel<CELL, ACT, RB, KPT><<<grid, threads, smem, st>>>(xg, U, mask, hout, T, B, H);
  } else {
    const size_t smem = sizeof(float) * (size_t)(G * RB * KP + KS * RB * CG);
    rnn_backward_reg_kernel<CELL, ACT, RB, KPT><<<grid, threads, smem, st>>>(xg, U, mask, hout, cst, dhout, T, B, H);
  }
  SEQREC_CHECK_LAUNCH();
  return 0;
}

template <int CELL, int ACT, int RB>
int dispatch_kpt(bool fwd, float* xg, const float* U, const uint8_t* mask, float* hout, float* cst,
                 const float* dhout, int T, int B, int H, cudaStream_t st) {
  const int kpt = (((H + KS - 1) / KS) + 7) / 8 * 8;   // K-slice length, multiple of 8
  switch (kpt) {
    case 8: return launch_pair<CELL, ACT, RB, 8>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    case 16: return launch_pair<CELL, ACT, RB, 16>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    case 24: return launch_pair<CELL, ACT, RB, 24>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    default: return launch_pair<CELL, ACT, RB, 32>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
  }
}

template <int CELL, int ACT>
int dispatch_rb(int rb, bool fwd, float* xg, const float* U, const uint8_t* mask, float* hout, float* cst,
                const float* dhout, int T, int B, int H, cudaStream_t st) {
  switch (rb) {
    case 1: return dispatch_kpt<CELL, ACT, 1>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    case 2: return dispatch_kpt<CELL, ACT, 2>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    default: return dispatch_kpt<CELL, ACT, 4>(fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
  }
}

template <int CELL>
int dispatch_act(int act, int rb, bool fwd, float* xg, const float* U, const uint8_t* mask, float* hout, float* cst,
                 const float* dhout, int T, int B, int H, cudaStream_t st) {
  switch (act) {
    case SEQREC_ACT_RELU:
      return dispatch_rb<CELL, SEQREC_ACT_RELU>(rb, fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    case SEQREC_ACT_TANH:
      return dispatch_rb<CELL, SEQREC_ACT_TANH>(rb, fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    case SEQREC_ACT_LINEAR:
      return dispatch_rb<CELL, SEQREC_ACT_LINEAR>(rb, fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
    default: return -1002;
  }
}

}  // namespace

// Is the register-resident scan applicable?  (GRU / SimpleRNN, H <= 128)
bool rnn_reg_applicable(int cell, int H) {
  return (cell == SEQREC_CELL_GRU || cell == SEQREC_CELL_SIMPLE) && H <= 128;
}

// rb: batch rows per CTA (1, 2 or 4).  U is the untransposed recurrent kernel (H, G*H) for BOTH directions.
int rnn_reg_launch(int cell, int act, int rb, bool fwd, float* xg, const float* U, const uint8_t* mask, float* hout,
                   float* cst, const float* dhout, int T, int B, int H, cudaStream_t st) {
  if (cell == SEQREC_CELL_GRU)
    return dispatch_act<SEQREC_CELL_GRU>(act, rb, fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
  return dispatch_act<SEQREC_CELL_SIMPLE>(act, rb, fwd, xg, U, mask, hout, cst, dhout, T, B, H, st);
}
