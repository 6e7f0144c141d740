"""world_size-2 gloo test (CPU) of the data-parallel host logic: shard bookkeeping and the weighted gradient
all-reduce reproduce the single-process full-batch gradient (gradients produced by the oracle here; on GPUs the
same buffers come from inr_grad_step)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import inr_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


NET = {"network_input_size": 64, "network_output_size": 2, "network_depth": 3, "network_width": 32}


def _flat_grad(sd, x, gt):
    tr = []
    out = O.siren_forward(sd, x, 3, trace=tr)
    _, dout = O.loss_l2(out, gt)
    grads, _ = O.siren_backward(sd, x, tr, dout, 3)
    return torch.cat([grads[k].reshape(-1) for k in sd])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_implicit_neural_representations_b200.parallel import allreduce_mean_, shard_rows
    torch.manual_seed(0)
    sd = O.siren_init(dict(NET))
    g = torch.Generator().manual_seed(1)
    n_rows, gbs = 1001, 300            # ragged: last global batch has 101 rows -> 51 + 50
    x = torch.randn(n_rows, 64, generator=g)
    gt = torch.rand(n_rows, 2, generator=g)
    res = []
    for start in range(0, n_rows, gbs):
        s, c = shard_rows(start, gbs, n_rows, rank, world)
        n_glob = min(start + gbs, n_rows) - start
        flat = _flat_grad(sd, x[s:s + c], gt[s:s + c])
        allreduce_mean_(flat, weight=c / n_glob * world)
        res.append(flat)
    if rank == 0:
        q.put([r.clone() for r in res])
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    sd = O.siren_init(dict(NET))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1001, 64, generator=g)
    gt = torch.rand(1001, 2, generator=g)
    for i, start in enumerate(range(0, 1001, 300)):
        full = _flat_grad(sd, x[start:start + 300], gt[start:start + 300])
        assert torch.allclose(got[i], full, rtol=1e-4, atol=1e-8), i


def test_shard_rows_partition():
    from mri_implicit_neural_representations_b200.parallel import shard_rows
    for n_rows, gbs, world in [(1001, 300, 2), (10000, 10000, 8), (17, 5, 4), (5, 8, 8)]:
        for start in range(0, n_rows, gbs):
            covered = []
            for r in range(world):
                s, c = shard_rows(start, gbs, n_rows, r, world)
                covered += list(range(s, s + c))
            assert covered == list(range(start, min(start + gbs, n_rows)))
