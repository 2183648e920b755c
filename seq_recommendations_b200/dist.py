"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in the CPU tests).

The path shards over SEQUENCES: no cross-sequence dependence exists anywhere in the forward or backward
pass, so each rank runs the whole hot path on B/P sequences.  What must be exchanged to make the step equal to the
single-process global-batch step:
  * n_valid = sum(mask) and loss_sum -- the Keras masked objective divides by the GLOBAL number of unmasked steps
  * the dense gradients (dU, db, dW_out, db_out): sum all-reduce of un-normalised-by-rank partials
  * dW_in: dense sum all-reduce when V*G*H is small, otherwise an all-gather of (ids, dxp rows) that every rank
    scatter-adds locally, so each replica applies the identical row-sparse update
The global gradient norm (clipnorm) is then computed identically on every rank without a further collective.
"""
import os

import torch
import torch.distributed as dist


class Comm:
    """Thin wrapper so the engine never touches torch.distributed directly (world_size 1 == no-ops)."""

    def __init__(self, group=None):
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    def all_reduce_sum(self, t, async_op=False):
        if not self.enabled:
            return None
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def all_reduce_max(self, t):
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def all_gather_cat(self, t):
        """Concatenate equally shaped per-rank tensors along dim 0."""
        if not self.enabled:
            return t
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def all_gather_cat_async(self, t):
        """Same, issued asynchronously: returns (out, handle); handle.wait() before `out` is read."""
        if not self.enabled:
            return t, None
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        h = dist.all_gather_into_tensor(out, t.contiguous(), group=self.group, async_op=True)
        return out, h

    def reduce_scatter_sum(self, out, inp):
        """out (n, ...) = sum over ranks of this rank's block of inp (world*n, ...)."""
        if not self.enabled:
            out.copy_(inp)
            return
        dist.reduce_scatter_tensor(out, inp.contiguous(), op=dist.ReduceOp.SUM, group=self.group)

    def all_to_all_rows(self, inp):
        """inp (world*n, ...): block j goes to rank j; returns (world*n, ...) whose block i came from rank i."""
        if not self.enabled:
            return inp
        out = torch.empty_like(inp)
        dist.all_to_all_single(out, inp.contiguous(), group=self.group)
        return out

    def broadcast(self, t, src=0):
        """In-place broadcast of rank `src`'s tensor (replicas must start from identical weights)."""
        if self.enabled:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                           group=self.group)
        return t

    def barrier(self):
        if self.enabled:
            dist.barrier(group=self.group)


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return Comm()
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return Comm()


def shard_rows(n_rows, rank, world):
    """Contiguous, balanced split of the batch: rank r owns rows [lo, hi)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def reduce_step_scalars(comm, n_valid, loss_sum):
    """Global n_valid and loss_sum (float tensors of one element each, reduced in place)."""
    comm.all_reduce_sum(n_valid)
    comm.all_reduce_sum(loss_sum)
    return n_valid, loss_sum


def embedding_grad_mode(V, GH, n_tokens_global):
    """'dense' all-reduce of dW_in (V*GH floats) vs 'rows' all-gather of (ids, dxp) (n_tokens*GH floats)."""
    return "dense" if V * GH <= n_tokens_global * GH else "rows"
