"""GPU parity of the fused multi-scale step (inr_train_step_dist / inr_grad_step_dist): MultiscaleBoundedFourier /
MultiscaleKFourier heads + per-head loss on the full target + 0.1 * ConsistencyLoss, exactly the loop body of the
reference's src/train_kspace_multiscale.py:164-192 (row mask on the per-head loss only, consistency on all rows), against
the ORACLE's forward + closed-form losses + torch autograd (oracle.loss_consistency / loss_logspace are pinned against the
reference's ConsistencyLoss / LogSpaceLoss in tests/test_losses_vs_reference.py)."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O

pytestmark = pytest.mark.gpu
PAIRS = [(0.0, 0.4), (0.0, 0.8), (0.0, 1.1), (0.0, 5.0)]
OPTS = {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.0}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _oracle_composite(P, x, dist, gt, mask, loss, bounded):
    bounds = [p for p in PAIRS for _ in (0, 1)] if bounded else None
    outs = O.multiscale_forward(P, x, 8, dist if bounded else None, bounds)
    total, douts = O.loss_consistency([o.detach() for o in outs], dist, PAIRS, 0.1)
    sel = mask if mask is not None else torch.ones(gt.shape[0], dtype=torch.bool)
    for k, o in enumerate(outs):
        if loss == "LSL":
            v, g = O.loss_logspace(o.detach()[sel], gt[sel], OPTS["hdr_eps"])
        else:
            v, g = O.LOSS_TRAIN[loss](o.detach()[sel], gt[sel])
        total = total + v
        full = torch.zeros_like(o)
        full[sel] = g
        douts[k] = douts[k] + full
    return outs, total, douts


@pytest.mark.parametrize("bounded,loss,masked", [(True, "LSL", False), (True, "LSL", True), (False, "L2", True)])
def test_fused_multiscale_loss_and_gradients_vs_oracle(inr, bounded, loss, masked):
    net = dict(G.NET_MFN)
    torch.manual_seed(17)
    encB = O.encoder_init(G.ENC_GAUSS)
    sd = O.multiscale_init(dict(net), bounded=bounded)
    bs = 700
    g = torch.Generator().manual_seed(4)
    coords = torch.rand(bs, 3, generator=g) * 2 - 1
    gt = torch.randn(bs, 2, generator=g) * 0.05
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    mask = ((torch.arange(bs) // 3) % 2 == 0) if masked else None
    pnet = dict(net)
    if bounded:
        pnet["boundaries"] = [p for p in PAIRS for _ in (0, 1)]
    plan = inr.Plan("BoundedFourier" if bounded else "MultiscaleFourier", pnet, G.ENC_GAUSS)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    x = O.encode(coords, encB, "gauss")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs, val, douts = _oracle_composite(P, x, dist, gt, mask, loss, bounded)
    grs = torch.autograd.grad(outs, list(P.values()), grad_outputs=douts, allow_unused=True)
    opts = dict(OPTS)
    opts["consistency"] = (PAIRS, 0.1)
    out = torch.empty(bs, 8, device="cuda")
    eng.grad_step(loss, coords.cuda(), gt.cuda(), bs, mask=None if mask is None else mask.to(torch.uint8).cuda(), loss_opts=opts,
                  out=out, dist=dist.cuda())
    for k, o in enumerate(outs):
        assert rel(out[:, 2 * k:2 * k + 2], o) <= 1e-3, k
    assert abs(float(eng.loss_out) - float(val)) <= 2e-3 * abs(float(val)), (float(eng.loss_out), float(val))
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for k, ref in zip(P.keys(), grs):
        if ref is None:
            assert float(gv[k].abs().max()) == 0.0, k
        else:
            # the loss gradient is evaluated on the engine's own fp16-operand outputs (LSL divides by (|x| + eps)^2): end to
            # end, not teacher-forced -- same band as the autograd-face composite test
            assert rel(gv[k], ref) <= 5e-3, (k, rel(gv[k], ref))


def test_fused_multiscale_steps_follow_the_oracle_loop(inr):
    """Four fused steps (grid-order batches, cursor on the device) of MultiscaleBoundedFourier + LSL + consistency against the
    oracle loop with torch.optim.Adam."""
    net = dict(G.NET_MFN)
    torch.manual_seed(23)
    encB = O.encoder_init(G.ENC_GAUSS)
    sd = O.multiscale_init(dict(net), bounded=True)
    n, bs, lr = 1200, 400, 2e-5
    g = torch.Generator().manual_seed(6)
    coords = torch.rand(n, 3, generator=g) * 2 - 1
    gt = torch.randn(n, 2, generator=g) * 0.05
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    pnet = dict(net)
    pnet["boundaries"] = [p for p in PAIRS for _ in (0, 1)]
    plan = inr.Plan("BoundedFourier", pnet, G.ENC_GAUSS)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=lr)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(P.values()), lr=lr)
    opts = dict(OPTS)
    opts["consistency"] = (PAIRS, 0.1)
    c, y, d = coords.cuda(), gt.cuda(), dist.cuda()
    eng.cursor.zero_()
    for step in range(4):
        i = (step * bs) % n
        outs, val, douts = _oracle_composite(P, O.encode(coords[i:i + bs], encB, "gauss"), dist[i:i + bs], gt[i:i + bs], None, "LSL", True)
        opt.zero_grad()
        torch.autograd.backward(outs, douts)
        opt.step()
        if step == 3:
            eng.cursor.zero_()
        eng.train_step("LSL", c, y, bs, loss_opts=opts, use_cursor=True, dist=d)
        assert abs(float(eng.loss_out) - float(val)) <= (2e-3 if step == 0 else 6e-3) * abs(float(val)), (step, float(eng.loss_out), float(val))
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        got = eng.params[off:off + rows * cols].cpu()
        assert abs(float(got.norm()) - float(P[k].norm())) <= 1e-3 * float(P[k].norm()), k


def test_fused_multiscale_full_size_batch_config4(inr):
    """BASELINE configs[3] size (MultiscaleBoundedFourier w512, 8 stages, 4 heads, LSL + 0.1 * consistency, batch 100 000 =
    782 row tiles, last one ragged; reference src/train_kspace_multiscale.py:164-192).  (a) every head, the loss and every
    gradient against the oracle composite on the whole batch, (b) size-independent properties: a row permutation of the
    batch (coordinates, targets, distances together) leaves every row's heads bit-identical and loss / gradients unchanged
    up to the order of the fixed-order reductions; the same step twice gives the same bits."""
    net = dict(G.NET_MFN)
    torch.manual_seed(31)
    encB = O.encoder_init(G.ENC_GAUSS)
    sd = O.multiscale_init(dict(net), bounded=True)
    bs = 100000
    g = torch.Generator().manual_seed(8)
    coords = torch.rand(bs, 3, generator=g) * 2 - 1
    gt = torch.randn(bs, 2, generator=g) * 0.05
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    pnet = dict(net)
    pnet["boundaries"] = [p for p in PAIRS for _ in (0, 1)]
    opts = dict(OPTS)
    opts["consistency"] = (PAIRS, 0.1)

    def grad_step(c, y, d):
        plan = inr.Plan("BoundedFourier", pnet, G.ENC_GAUSS)
        eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
        eng.load_tensors(list(sd.values()))
        eng.set_encoder(encB)
        out = torch.empty(bs, 8, device="cuda")
        eng.grad_step("LSL", c.cuda(), y.cuda(), bs, loss_opts=opts, out=out, dist=d.cuda())
        torch.cuda.synchronize()
        return eng, out.cpu(), eng.grads.clone(), float(eng.loss_out)

    eng, out, grads, loss = grad_step(coords, gt, dist)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs, val, douts = _oracle_composite(P, O.encode(coords, encB, "gauss"), dist, gt, None, "LSL", True)
    grs = torch.autograd.grad(outs, list(P.values()), grad_outputs=douts, allow_unused=True)
    for k, o in enumerate(outs):
        assert rel(out[:, 2 * k:2 * k + 2], o) <= 1e-3, k
    assert abs(loss - float(val)) <= 2e-3 * abs(float(val)), (loss, float(val))
    gv = dict(zip(sd.keys(), eng._views(grads)))
    for k, ref in zip(P.keys(), grs):
        if ref is None:
            assert float(gv[k].abs().max()) == 0.0, k
        else:
            assert rel(gv[k], ref) <= 5e-3, (k, rel(gv[k], ref))       # end to end, same band as the 700-row case above
    _, out2, grads2, loss2 = grad_step(coords, gt, dist)
    assert torch.equal(out2, out) and torch.equal(grads2, grads) and loss2 == loss
    perm = torch.randperm(bs, generator=g)
    _, out_p, grads_p, loss_p = grad_step(coords[perm], gt[perm], dist[perm])
    assert torch.equal(out_p, out[perm])
    assert abs(loss_p - loss) <= 1e-5 * abs(loss)
    assert rel(grads_p, grads) <= 1e-4
