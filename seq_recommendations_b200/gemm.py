"""Plain products of the path: C (+)= op(A) . op(B) (+ bias) on contiguous 2-D fp32 device tensors.

Large products run on the tcgen05 GEMM (csrc/gemm_tc.cu: K-major bf16 hi/lo operands staged by seqrec_split_bf16,
3-pass split = fp32-grade, or one pass in bf16 mode); products too small to fill 128 x 128 tiles run on the fp32 SIMT
kernels (csrc/gemm.cu, csrc/dense_ops.cu).  Used for K2 -- the dense half of the RNN input projection
(/root/reference/model.py:245-255, :354-358) -- and for the logit terms of the history-feature branches and their
gradients (model.py:376-392)."""
import torch

from ._lib import call, ptr


def gemm(owner, A, Bm, C, form, bias=None, accumulate=False):
    """form: 'nn' A(M,K).B(K,N); 'nt' A(M,K).B(N,K)^T; 'tn' A(K,M)^T.B(K,N) ('tn' always accumulates into C, which holds
    zeros or a partial sum: it is the weight-gradient form).  owner supplies device, stream, tc_mode, tc_x3."""
    st = owner.stream
    if form == "nn":
        M, K = A.shape
        N = Bm.shape[1]
    elif form == "nt":
        M, K = A.shape
        N = Bm.shape[0]
    else:
        K, M = A.shape
        N = Bm.shape[1]
    if owner.tc_mode != "off" and M >= 128 and N >= 64 and K >= 32 and M * N * K >= (1 << 24):
        Kp = (K + 63) // 64 * 64
        bf = torch.bfloat16
        x3 = owner.tc_x3

        def stage(src, transpose, rows_out):
            hi = torch.zeros((rows_out, Kp), dtype=bf, device=owner.device)
            lo = torch.zeros((rows_out, Kp), dtype=bf, device=owner.device) if x3 else None
            call("seqrec_split_bf16", ptr(src), None, ptr(hi), ptr(lo), src.shape[0], src.shape[1], Kp,
                 1 if transpose else 0, st)
            return hi, lo
        a_hi, a_lo = stage(A, form == "tn", M)
        b_hi, b_lo = stage(Bm, form != "nt", N)
        call("seqrec_gemm_tc", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), ptr(bias), ptr(C), M, N, K, Kp, Kp, N,
             1 if (accumulate or form == "tn") else 0, 1 if x3 else 0, st)
        return
    if form == "nn":
        call("seqrec_gemm_nn", ptr(A), ptr(Bm), ptr(bias), ptr(C), M, N, K, 1 if accumulate else 0, st)
    elif form == "nt":
        call("seqrec_gemm_nt", ptr(A), ptr(Bm), ptr(C), M, N, K, K, K, N, 1 if accumulate else 0, st)
        if bias is not None:
            call("seqrec_add_rows", ptr(C), None, ptr(bias), None, M, N, st)
    else:
        call("seqrec_gemm_tn_atomic", ptr(A), ptr(Bm), ptr(C), M, N, K, st)
