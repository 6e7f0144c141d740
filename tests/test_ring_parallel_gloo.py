"""Ring models, one per rank (reference src/train_variations/train_clustering.py): world_size-2 gloo
test (CPU) of the host logic -- ring -> rank assignment, the shared jitter stream, and the validation assembly
(disjoint writer masks + one all-reduce) against the reference's sequential `batch_rec[ind] = out_i` overwrite."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

RADII = np.array([0.0, 0.25, 0.5, 0.9, 5.0])
NO_MODELS = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ring_output(i, coords):
    """Stand-in for ring model i's prediction: any deterministic function that differs per ring."""
    return torch.stack([coords[:, 1] * (i + 1) + 0.1 * i, coords[:, 2] - i], dim=1)


def _data():
    g = torch.Generator().manual_seed(3)
    coords = torch.rand(4000, 3, generator=g) * 2 - 1
    # put some rows exactly on ring edges (they belong to two rings; the outer one must win)
    coords[:50, 1] = 0.25
    coords[:50, 2] = 0.0
    coords[50:80, 1] = 0.0
    coords[50:80, 2] = 0.5
    return coords, torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)


def _reference_assembly(coords, d):
    rec = torch.zeros(coords.shape[0], 2)
    for i in range(NO_MODELS):                      # reference :218-232
        ind = torch.where((d >= RADII[i]) & (d <= RADII[i + 1]))
        if ind[0].numel():
            rec[ind] = _ring_output(i, coords[ind])
    return rec


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_implicit_neural_representations_b200 import parallel as P
    coords, d = _data()
    mine = P.owned_rings(NO_MODELS, rank, world)
    rng = np.random.RandomState(0)
    limits = [P.ring_jitter(rng, RADII, NO_MODELS) for _ in range(5)]
    writers = P.ring_writer_masks(d, RADII, NO_MODELS)
    partial = torch.zeros(coords.shape[0], 2)
    for i in mine:
        rows = torch.nonzero(writers[i]).squeeze(1)
        partial[rows] = _ring_output(i, coords[rows])
    rec = P.combine_ring_outputs(partial)
    q.put((rank, mine, limits, rec.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_ring_models_over_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get() for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2] and got[1][1] == [1, 3]
    assert got[0][2] == got[1][2]                                    # same jitter stream on every rank
    coords, d = _data()
    want = _reference_assembly(coords, d)
    for _, _, _, rec in got:
        assert torch.equal(rec, want)


def test_ring_helpers_single_process():
    from mri_implicit_neural_representations_b200 import parallel as P
    for world in (1, 2, 4, 8):
        owned = [P.owned_rings(NO_MODELS, r, world) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(NO_MODELS))
        assert max(len(o) for o in owned) == -(-NO_MODELS // world)
    coords, d = _data()
    writers = P.ring_writer_masks(d, RADII, NO_MODELS)
    total = torch.stack(writers).sum(0)
    assert int(total.max()) == 1 and int(total.min()) == 1              # every row written by exactly one ring
    assert bool(writers[1][:50].all()) and not bool(writers[0][:50].any())     # edge rows go to the outer ring
    rng = np.random.RandomState(1)
    for r0, r1 in P.ring_jitter(rng, RADII, NO_MODELS):
        assert r0 >= 0.0 and r1 > r0
    lim = P.ring_jitter(np.random.RandomState(1), RADII, NO_MODELS)
    ref_rng = np.random.RandomState(1)
    for i in range(NO_MODELS):                                          # reference :175-176, draw for draw
        r0 = max(0, RADII[i] - np.abs(ref_rng.normal(0, 0.05)))
        r1 = RADII[i + 1] + np.abs(ref_rng.normal(0, 0.05))
        assert lim[i] == (float(r0), float(r1))
    # combine without a process group is the identity
    x = torch.ones(3, 2)
    assert P.combine_ring_outputs(x) is x
