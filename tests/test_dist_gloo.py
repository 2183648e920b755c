"""world_size-2 gloo tests (CPU) of the data-parallel arithmetic in seq_recommendations_b200/dist.py: sharding by
sequences, GLOBAL n_valid normalisation, sum all-reduce of gradients and the two embedding-gradient exchange modes give
exactly the single-process global-batch step.  The per-shard compute is the oracle (tests may use it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import keras_semantics as ks
from seq_recommendations_b200 import dist, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cell, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    comm = dist.init_from_env("gloo")
    assert comm.enabled and comm.world == world and comm.rank == rank
    V, H, T, B = 23, 6, 7, 9
    ws = synthetic.make_weights(cell, V, H, seed=1)
    ids, tgt = synthetic.make_batch(V, T, B, seed=2, min_len=1)
    lo, hi = dist.shard_rows(B, rank, world)
    mod = ks.Model(cell, "tanh", ws, dtype=torch.float64)
    i, t = torch.tensor(ids[lo:hi].astype(np.int64)), torch.tensor(tgt[lo:hi].astype(np.int64))
    m = i >= 0
    # what one rank's kernels produce: un-normalised loss sum and gradients scaled by 1/n_valid_GLOBAL
    n_valid = m.sum().to(torch.float64).reshape(1)
    ps = [p.clone().requires_grad_(True) for p in mod.params()]
    mod.set_params(ps)
    _, ce, _ = mod.loss(i, t, m)
    loss_sum = ce.sum().reshape(1)
    n_glob, _ = dist.reduce_step_scalars(comm, n_valid.clone(), loss_sum.detach().clone())
    grads = torch.autograd.grad(loss_sum[0] / n_glob[0], ps)
    loss_glob = loss_sum.detach().clone()
    comm.all_reduce_sum(loss_glob)
    dense = [g.clone() for g in grads[1:]]
    flat = torch.cat([g.reshape(-1) for g in dense])
    comm.all_reduce_sum(flat)
    # embedding gradient, 'dense' mode: all-reduce the (V, GH) table gradient
    dW_dense = grads[0].clone()
    comm.all_reduce_sum(dW_dense)
    # 'rows' mode: all-gather (ids, dxp rows) and scatter-add locally.  dxp rows are recovered from the oracle by
    # differentiating w.r.t. the gathered input projection.
    xp = ks.input_projection(ps[0].detach(), ps[2].detach(), ids=i, mask=m).requires_grad_(True)
    Hh = ks.rnn_forward(xp, ps[1].detach(), m, cell, "tanh")
    _, ce2, _ = ks.masked_loss(ks.logits(Hh, ps[3].detach()), t, m)
    dxp, = torch.autograd.grad(ce2.sum() / n_glob[0], xp)
    pad = (B + world - 1) // world - (hi - lo)             # equal shapes for all_gather: pad the short shard
    ids_flat = torch.cat([i.reshape(-1), torch.full((pad * T,), -1, dtype=torch.int64)])
    dxp_flat = torch.cat([dxp.reshape(-1, dxp.shape[-1]), torch.zeros(pad * T, dxp.shape[-1], dtype=dxp.dtype)])
    all_ids = comm.all_gather_cat(ids_flat)
    all_dxp = comm.all_gather_cat(dxp_flat)
    dW_rows = torch.zeros_like(dW_dense)
    keep = all_ids >= 0
    dW_rows.index_add_(0, all_ids[keep], all_dxp[keep])
    if rank == 0:
        torch.save(dict(loss=float(loss_glob[0] / n_glob[0]), flat=flat, dW_dense=dW_dense, dW_rows=dW_rows,
                        n=float(n_glob[0])), out)
    comm.barrier()
    tdist.destroy_process_group()


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
def test_two_rank_step_equals_global_batch_step(tmp_path, cell):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), cell, out), nprocs=2, join=True)
    got = torch.load(out)
    V, H, T, B = 23, 6, 7, 9
    ws = synthetic.make_weights(cell, V, H, seed=1)
    ids, tgt = synthetic.make_batch(V, T, B, seed=2, min_len=1)
    mod = ks.Model(cell, "tanh", ws, dtype=torch.float64)
    i, t = torch.tensor(ids.astype(np.int64)), torch.tensor(tgt.astype(np.int64))
    loss, gs = mod.grads(i, t, i >= 0)
    assert got["n"] == float((i >= 0).sum())
    assert abs(got["loss"] - float(loss)) < 1e-12
    assert torch.allclose(got["flat"], torch.cat([g.reshape(-1) for g in gs[1:]]), atol=1e-12)
    assert torch.allclose(got["dW_dense"], gs[0], atol=1e-12)
    assert torch.allclose(got["dW_rows"], gs[0], atol=1e-12)


def test_shard_rows_and_mode_choice():
    for n in (1, 7, 8, 256, 1023):
        for world in (1, 2, 3, 8):
            spans = [dist.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert dist.embedding_grad_mode(10000, 384, 8 * 12800) == "dense"
    assert dist.embedding_grad_mode(1000000, 768, 8 * 6400) == "rows"
    c = dist.Comm()
    assert not c.enabled and c.world == 1 and c.all_gather_cat(torch.ones(2)).shape == (2,)
