"""Two-GPU test of the optimiser-fused data-parallel gradient exchange (inr_adam_step_peers): replicas stay bit-identical
across ranks and match the NCCL all-reduce(avg) path.  Skipped on single-GPU boxes."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200.parallel import PeerGradExchange, allreduce_mean_
    from oracle import golden_util as G
    from oracle.cases import case_setup
    res = {}
    for case, model in (("siren_l2", "SIREN"), ("wire_l2", "WIRE")):
        model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(case)
        bs = 256
        g = torch.Generator().manual_seed(100 + rank)                 # every rank trains on its own rows
        idx = torch.randperm(coords.shape[0], generator=g)[:bs]
        c, y = coords[idx].to(dev), gt[idx].to(dev)
        finals = []
        for mech in ("peer", "nccl"):
            plan = inr.Plan(model, net, enc_cfg)
            eng = inr.ChainEngine(plan, max_batch=bs, device=dev, lr=G.LR)
            eng.load_tensors(list(sd.values()))
            eng.set_encoder(encB)
            ex = PeerGradExchange(plan.n_params, dev) if mech == "peer" else None
            for step in range(6):
                if ex is not None:
                    eng.grad_step(loss_kind, c, y, bs, loss_opts=opts, grads=ex.grads(step & 1))
                    eng.adam_step_peers(ex, step & 1)
                else:
                    eng.grad_step(loss_kind, c, y, bs, loss_opts=opts)
                    allreduce_mean_(eng.grads)
                    eng.adam_step()
            torch.cuda.synchronize()
            finals.append(eng.params.clone())
        gathered = [torch.empty_like(finals[0]) for _ in range(world)]
        dist.all_gather(gathered, finals[0])
        res[case] = (all(torch.equal(gathered[0], t) for t in gathered),
                     float((finals[0] - finals[1]).abs().max()), float(finals[1].abs().max()))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_exchange_matches_nccl_and_keeps_replicas_identical():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = q.get()
    for case, (same_across_ranks, max_diff, scale) in res.items():
        assert same_across_ranks, case
        # mean over 2 ranks: (a + b) * 0.5 is exact either way; allow one ulp-level difference from NCCL's reduction order
        assert max_diff <= 1e-6 * scale, (case, max_diff, scale)
