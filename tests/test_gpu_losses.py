"""GPU parity of every fused loss (value + full parameter gradient through the SIREN chain), the row mask,
the regulariser gradients and the gradient-only step, against the oracle's closed forms (which are pinned to the
reference's loss classes by tests/golden/losses.json)."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup

pytestmark = pytest.mark.gpu
TOL = 1e-3
HDR = {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _setup(inr, bs, kspace_like):
    model_kind, net, enc_cfg, _, _, sd, encB, coords, gt, _ = case_setup("siren_l2")
    coords = coords[:bs]
    g = torch.Generator().manual_seed(99)
    gt = (torch.randn(bs, 2, generator=g) * 0.05) if kspace_like else (torch.rand(bs, 2, generator=g) * 0.8 + 0.1)
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=0.0)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.siren_forward(sd, x, 4, trace=tr)
    return plan, eng, sd, coords, gt, x, tr, out_ref


@pytest.mark.parametrize("loss,masked", [("L1", False), ("MSLE", False), ("tanh", False), ("LSL", False), ("LSL", True),
                                         ("HDR", False), ("HDR", True), ("L2", True)])
def test_fused_loss_value_and_gradients(inr, loss, masked):
    bs = 900
    plan, eng, sd, coords, gt, x, tr, out_ref = _setup(inr, bs, kspace_like=loss in ("HDR", "LSL"))
    if loss == "MSLE":
        gt = gt.abs()
    mask = (torch.arange(bs) % 2 == 0) if masked else torch.ones(bs, dtype=torch.bool)
    out_dev = torch.empty(bs, 2, device="cuda")
    g = eng.grad_step(loss, coords.cuda(), gt.cuda(), bs, mask=mask.to(torch.uint8).cuda() if masked else None,
                      loss_opts=HDR, out=out_dev)
    torch.cuda.synchronize()
    assert rel(out_dev, out_ref) <= TOL
    # The loss stage is judged teacher-forced on the engine's own network output: L1's sign(e) is discontinuous and
    # HDR's 2 log(|e|/d) e/|e|^2 has a 1/|e|^2 sensitivity, so the 1.5e-4 forward error would otherwise dominate.
    o_in = out_dev.cpu() if loss in ("L1", "HDR") else out_ref
    o_sel, g_sel = o_in[mask], gt[mask]
    if loss == "HDR":
        val, dsel, _ = O.loss_hdr(o_sel, g_sel, coords, HDR["hdr_ff_sigma"], HDR["hdr_eps"], HDR["hdr_ff_factor"])
    elif loss == "LSL":
        val, dsel = O.loss_logspace(o_sel, g_sel, HDR["hdr_eps"])
    else:
        val, dsel = O.LOSS_TRAIN[loss](o_sel, g_sel)
    if loss == "MSLE" and not torch.isfinite(val):
        pytest.skip("MSLE undefined for this draw (log of a negative prediction), as in the reference")
    dout = torch.zeros_like(out_ref)
    dout[mask] = dsel
    grads_ref, _ = O.siren_backward(sd, x, tr, dout, 4)
    assert abs(float(eng.loss_out) - float(val)) <= TOL * abs(float(val)), (float(eng.loss_out), float(val))
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = g[off:off + rows * cols].view(grads_ref[k].shape)
        assert rel(gv, grads_ref[k]) <= TOL, (k, rel(gv, grads_ref[k]))


@pytest.mark.parametrize("kind", ["L1", "L2"])
def test_regulariser_gradient_in_adam(inr, kind):
    """reg folded into the optimiser kernel: first Adam moment after one step = (1-b1) * (g + dReg/dp)."""
    bs, lam = 640, 1e-3
    plan, eng, sd, coords, gt, x, tr, out_ref = _setup(inr, bs, kspace_like=False)
    eng.hyper[5 if kind == "L1" else 6] = lam
    val, dout = O.loss_l2(out_ref, gt)
    grads_ref, _ = O.siren_backward(sd, x, tr, dout, 4)
    eng.train_step("L2", coords.cuda(), gt.cuda(), bs)
    torch.cuda.synchronize()
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        p = sd[k]
        reg_g = lam * torch.sign(p) if kind == "L1" else 2 * lam * p
        gv = eng.exp_avg[off:off + rows * cols].view(p.shape) / 0.1
        assert rel(gv, grads_ref[k] + reg_g) <= TOL, k


def test_grad_step_then_adam_equals_fused_step(inr):
    """data-parallel building blocks: inr_grad_step + inr_adam_step == inr_train_step (bit-identical)."""
    bs = 512
    plan, e1, sd, coords, gt, *_ = _setup(inr, bs, kspace_like=False)
    _, e2, *_ = _setup(inr, bs, kspace_like=False)
    for e in (e1, e2):
        e.set_lr(5e-4)
    c, y = coords.cuda(), gt.cuda()
    for _ in range(2):
        e1.train_step("L2", c, y, bs)
        e2.grad_step("L2", c, y, bs)
        e2.adam_step()
    torch.cuda.synchronize()
    assert torch.equal(e1.params, e2.params)
    assert torch.equal(e1.wpack, e2.wpack)


def test_fused_step_is_bit_reproducible(inr):
    bs = 1000
    outs = []
    for _ in range(2):
        plan, eng, sd, coords, gt, *_ = _setup(inr, bs, kspace_like=False)
        eng.set_lr(5e-4)
        for _ in range(3):
            eng.train_step("L2", coords.cuda(), gt.cuda(), bs)
        torch.cuda.synchronize()
        outs.append(eng.params.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("loss,weight", [("tanh", 0.5), ("L2", 1e-4)])
def test_fused_tv_term_per_coil_batch(inr, loss, weight):
    """Per-coil loop of the reference (src/train.py:172-182): TV on out.view(H, W, 2) over ALL rows, then the row mask,
    then the main loss.  The TV gradient is sign-based, so it is judged teacher-forced on the engine's own output."""
    H, W = 24, 32
    bs = H * W
    plan, eng, sd, coords, gt, x, tr, out_ref = _setup(inr, bs, kspace_like=False)
    mask = ((torch.arange(bs) // W) % 2 == 0)                    # grid-2*1: every other image row is sampled
    out_dev = torch.empty(bs, 2, device="cuda")
    g = eng.grad_step(loss, coords.cuda(), gt.cuda(), bs, mask=mask.to(torch.uint8).cuda(),
                      loss_opts={"tv": (H, W, weight)}, out=out_dev)
    torch.cuda.synchronize()
    assert rel(out_dev, out_ref) <= TOL
    val_tv, g_tv = O.loss_tv(out_dev.cpu(), H, W, weight)
    val, dsel = O.LOSS_TRAIN[loss](out_ref[mask], gt[mask])
    dout = g_tv.clone()
    dout[mask] += dsel
    assert float(g_tv.norm()) > (0.05 if weight > 0.1 else 0.0) * float(dsel.norm())     # the TV term matters in this test
    grads_ref, _ = O.siren_backward(sd, x, tr, dout, 4)
    assert abs(float(eng.loss_out) - float(val + val_tv)) <= TOL * abs(float(val + val_tv))
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = g[off:off + rows * cols].view(grads_ref[k].shape)
        assert rel(gv, grads_ref[k]) <= TOL, (k, rel(gv, grads_ref[k]))
    # shape check of the ABI: TV needs one whole coil per batch
    with pytest.raises(Exception):
        eng.grad_step(loss, coords.cuda(), gt.cuda(), bs - W, loss_opts={"tv": (H, W, weight)}, out=out_dev)
