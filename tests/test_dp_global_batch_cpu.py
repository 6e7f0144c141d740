"""CPU test of the data-parallel loss normalisation (trainer.dp_batch_table + parallel.shard_rows): the mean over ranks of
the per-rank losses / output gradients, each normalised with the table's share of the GLOBAL batch, equals the loss /
gradient of the whole batch in one process -- for L2 and for HDR (product of two batch means: the filter mean over all
rows times the mean over the masked rows, reference src/metrics/losses.py:241-262), with a row mask and a ragged last
batch.  The per-rank arithmetic is the oracle's; on the GPU the same table feeds inr_loss_desc.dp_norm."""
import torch

from oracle import inr_oracle as O
from mri_implicit_neural_representations_b200.parallel import shard_rows
from mri_implicit_neural_representations_b200.trainer import dp_batch_table

OPTS = {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}


def _rank_pieces(loss, out, gt, coords, mask_rows, norm):
    """What one rank's kernels compute on its shard: numerators over the rows that enter the loss, divided by the
    table's count share; HDR uses the table's (global) filter mean."""
    m_share, fmean = float(norm[0]), float(norm[1])
    x, y = out[mask_rows].double(), gt[mask_rows].double()
    if loss == "L2":
        e = x - y
        val = (e * e).sum() * 0.5 / (m_share * 2)
        g = torch.zeros_like(out, dtype=torch.float64)
        g[mask_rows] = e / (m_share * 2)
        return val, g
    e = x - y
    ae2 = (e * e).sum(1)
    ax2 = (x * x).sum(1)
    d = ax2.sqrt() + OPTS["hdr_eps"]
    lg = torch.log(ae2.sqrt() / d)
    val = (lg * lg).sum() / m_share + OPTS["hdr_ff_factor"] * fmean * (ax2 / d ** 2).sum() / m_share
    g = torch.zeros_like(out, dtype=torch.float64)
    g[mask_rows] = (2 * lg / ae2)[:, None] * e / m_share + OPTS["hdr_ff_factor"] * fmean * 2 * x / (d ** 2)[:, None] / m_share
    return val, g


def test_mean_over_ranks_is_the_global_batch_loss_and_gradient():
    g = torch.Generator().manual_seed(3)
    n, gbs = 1030, 256                       # last global batch: 6 rows -> 2 + 2 + 1 + 1 over 4 ranks
    coords = torch.rand(n, 3, generator=g) * 2 - 1
    out = torch.randn(n, 2, generator=g) * 0.3
    gt = torch.randn(n, 2, generator=g) * 0.3
    mask = (torch.arange(n) // 7) % 2 == 0
    for world in (2, 4):
        for loss in ("L2", "HDR"):
            table = dp_batch_table(coords, mask.to(torch.uint8), gbs, world, loss, OPTS)
            assert table.shape == ((n + gbs - 1) // gbs, 2)
            for b, start in enumerate(range(0, n, gbs)):
                rows = slice(start, min(start + gbs, n))
                o, y, c, mk = out[rows], gt[rows], coords[rows], mask[rows]
                # single process, whole batch: the oracle's own loss (reference semantics)
                if loss == "L2":
                    v_ref, g_sel = O.loss_l2(o[mk].double(), y[mk].double())
                else:
                    v_ref, g_sel, _ = O.loss_hdr(o[mk].double(), y[mk].double(), c.double(), OPTS["hdr_ff_sigma"], OPTS["hdr_eps"],
                                                 OPTS["hdr_ff_factor"])
                g_ref = torch.zeros(o.shape, dtype=torch.float64)
                g_ref[mk] = g_sel
                vals, grads = [], torch.zeros(o.shape, dtype=torch.float64)
                covered = 0
                for r in range(world):
                    s, cnt = shard_rows(start, gbs, n, r, world)
                    covered += cnt
                    loc = slice(s - start, s - start + cnt)
                    v, gr = _rank_pieces(loss, o[loc], y[loc], c[loc], mk[loc], table[b])
                    vals.append(v)
                    grads[loc] = gr
                assert covered == o.shape[0]
                # mean over ranks of the per-rank values; per-rank gradients touch disjoint rows, so the rank-mean of the
                # PARAMETER gradient is (1 / world) * sum_r J_r^T g_r = J^T (grads / world)
                assert abs(float(sum(vals)) / world - float(v_ref)) <= 1e-6 * max(abs(float(v_ref)), 1e-12), (world, loss, b)   # the table is fp32
                assert float((grads / world - g_ref).abs().max()) <= 1e-6 * float(g_ref.abs().max()), (world, loss, b)


def test_plan_fits_spreads_samples_or_goes_data_parallel():
    import os
    import sys
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "src")
    sys.path.insert(0, src)
    try:
        import train
    finally:
        sys.path.remove(src)
    fits = [(f"s{i}", "file", i) for i in range(8)]
    assert train.plan_fits(fits, 0, 1) == ("single", fits)
    mode, mine = train.plan_fits(fits, 3, 4)
    assert mode == "independent" and mine == [fits[3], fits[7]]
    assert sorted(f for r in range(4) for f in train.plan_fits(fits, r, 4)[1]) == sorted(fits)
    assert train.plan_fits(fits[:1], 2, 4) == ("dp", fits[:1])
    assert train.plan_fits(fits, 1, 2, "dp") == ("dp", fits)
