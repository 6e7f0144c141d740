"""GPU parity tests of the multiplicative-filter-network path (FourierNet, MultiscaleKFourier) through the C ABI."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup, loss_and_grad

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _fourier(inr):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("fourier_l2")
    plan = inr.Plan("Fourier", net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    return plan, eng, net, loss_kind, opts, sd, encB, coords, gt


def test_fourier_forward_per_stage(inr):
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _fourier(inr)
    L, bs = net["network_depth"], coords.shape[0]
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.mfn_forward(sd, x, L, False, trace=tr)
    out = eng.forward(coords.cuda(), train=True)
    for i in range(L + 1):
        assert rel(eng.read_mfn_image("z", i, bs)[:bs], tr[i]) <= TOL, f"z{i}"
    assert rel(out, out_ref) <= TOL
    # teacher-forced stage: z_i recomputed in fp64 from the engine's own z_{i-1}
    x64 = x.double()
    for i in (1, L):
        zin = eng.read_mfn_image("z", i - 1, bs)[:bs].cpu().double()
        p = x64 @ sd[f"filters.{i}.linear.weight"].double().t() + sd[f"filters.{i}.linear.bias"].double()
        h = zin @ sd[f"linear.{i-1}.weight"].double().t() + sd[f"linear.{i-1}.bias"].double()
        assert rel(eng.read_mfn_image("z", i, bs)[:bs], torch.sin(p) * h) <= 6e-4, i


def test_fourier_gradients_and_fused_steps_vs_reference_golden(inr):
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _fourier(inr)
    L, bs = net["network_depth"], coords.shape[0]
    x = O.encode(coords, encB, "gauss")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.mfn_forward(P, x, L, False)
    val, dout = loss_and_grad(loss_kind, opts, o.detach(), gt, coords)
    gr = dict(zip(P.keys(), torch.autograd.grad(o, list(P.values()), grad_outputs=dout)))
    g = eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
    assert abs(float(eng.loss_out) - float(val)) <= 1e-5 * float(val)
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for k in sd:
        # end to end through 9 stages (not teacher-forced): fp16-operand rounding accumulates to ~1e-3 at the first filters
        assert rel(gv[k], gr[k]) <= 1.5e-3, (k, rel(gv[k], gr[k]))
    gold = G.load_golden("fourier_l2")
    for step in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
        assert abs(float(eng.loss_out) - gold["losses"][step]) <= 2e-4 * gold["losses"][step], step
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        dg = G.tensor_digest(eng.params[off:off + rows * cols].cpu())
        assert abs(dg["l2"] - gold["final"][k]["l2"]) <= 1e-4 * gold["final"][k]["l2"], k


def test_multiscale_heads_gradients_and_dead_parameters(inr):
    """MultiscaleKFourier through the autograd-face entry points: 4 heads out, external dL/dout per head in; parameters the
    reference never reaches (stage 8, unused heads: grad None there) receive exactly zero and are skipped by Adam."""
    model_kind, net, enc_cfg, loss_kind, opts, _, encB, coords, gt, mask = case_setup("fourier_l2")
    L, bs = net["network_depth"], coords.shape[0]
    torch.manual_seed(5)
    sd = O.multiscale_init(dict(net), bounded=False)
    plan = inr.Plan("MultiscaleFourier", net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    x = O.encode(coords, encB, "gauss")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs = O.multiscale_forward(P, x, L)
    out = eng.forward(coords.cuda(), train=True)
    assert out.shape == (bs, 8)
    for k, oo in enumerate(outs):
        assert rel(out[:, 2 * k:2 * k + 2], oo) <= TOL, k
    g0 = torch.Generator().manual_seed(3)
    douts = [torch.randn(bs, 2, generator=g0) * 1e-3 for _ in outs]
    grs = torch.autograd.grad(outs, list(P.values()), grad_outputs=douts, allow_unused=True)
    eng.backward(torch.cat(douts, dim=1).cuda())
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    before = eng.params.clone()
    n_dead = 0
    for k, ref in zip(P.keys(), grs):
        if ref is None:
            assert float(gv[k].abs().max()) == 0.0, k
            n_dead += 1
        else:
            # random dL/dy: the head gradient sum_rows dy * z has no error averaging (both factors fp16-rounded, the
            # engine's own z carries 4e-4): worst case of the 1-pass error model, judged end to end at 1.5e-3
            assert rel(gv[k], ref) <= 1.5e-3, (k, rel(gv[k], ref))
    assert n_dead == 14        # 14 tensors never get gradients (stage 8 + five unused heads)
    eng.adam_step()
    after = dict(zip(sd.keys(), eng.param_views()))
    for k, ref in zip(P.keys(), grs):
        moved = not torch.equal(after[k].cpu(), sd[k])
        assert moved == (ref is not None), k


def test_multiscale_module_composite_loss_vs_torch(inr):
    """Drop-in MultiscaleKFourier module inside the reference's composite multi-scale loss (per-head LSL on the full target
    + 0.1 * ConsistencyLoss, src/train_kspace_multiscale.py:173-190): every parameter gradient vs plain torch autograd."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "src"))
    from metrics.losses import ConsistencyLoss, LogSpaceLoss
    from mri_implicit_neural_representations_b200.modules import MultiscaleKFourier, Positional_Encoder
    torch.manual_seed(41)
    enc = Positional_Encoder(G.ENC_GAUSS, device="cuda")
    model = MultiscaleKFourier(dict(G.NET_MFN)).to("cuda")
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    bs = 500
    coords = torch.rand(bs, 3) * 2 - 1
    gt = torch.randn(bs, 2) * 0.05
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    pairs = [(0, 0.4), (0, 0.8), (0, 1.1), (0, 5)]
    lsl, cons = LogSpaceLoss({"hdr_ff_sigma": 1, "hdr_eps": 1e-2, "hdr_ff_factor": 0}), ConsistencyLoss(pairs)

    def composite(outs, gt_, dist_):
        loss = 0.1 * cons(outs, dist_)
        for o in outs:
            loss = loss + 0.5 * lsl(o.contiguous(), gt_)
        return loss

    outs = model(coords=enc.embedding(coords.cuda()), dist_to_center=dist.cuda())
    assert isinstance(outs, list) and len(outs) == 4
    loss = composite(outs, gt.cuda(), dist.cuda())
    loss.backward()
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = O.encode(coords, enc.B.cpu(), "gauss")
    routs = O.multiscale_forward(P, x, 8)
    rloss = composite(routs, gt, dist)
    rloss.backward()
    assert abs(float(loss) - float(rloss)) <= 2e-3 * abs(float(rloss))
    named = dict(model.named_parameters())
    for k in P:
        if P[k].grad is None:
            assert named[k].grad is None, k
        else:
            # LSL's 1/(|x|+eps)^2 weighting is evaluated on the engine's own (fp16-operand) outputs: not teacher-forced
            assert rel(named[k].grad, P[k].grad) <= 5e-3, (k, rel(named[k].grad, P[k].grad))


def test_bounded_fourier_forward_and_gradients(inr):
    """MultiscaleBoundedFourier (BASELINE config 4's model): BoundedLinear row masks from a 1-D dist_to_center, heads and
    every parameter gradient vs torch autograd on the oracle (reference mfn.py:281-286,344-356)."""
    from mri_implicit_neural_representations_b200.modules import MultiscaleBoundedFourier, Positional_Encoder
    torch.manual_seed(43)
    enc = Positional_Encoder(G.ENC_GAUSS, device="cuda")
    radii = [(0, 0.45), (0, 0.8), (0, 1.1), (0, 5)]
    bounds = [p for p in radii for _ in (0, 1)]
    model = MultiscaleBoundedFourier(dict(G.NET_MFN), boundaries=bounds).to("cuda")
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    assert "linear.0.linear.weight" in sd
    bs = 700
    coords = torch.rand(bs, 3) * 2 - 1
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    outs = model(coords=enc.embedding(coords.cuda()), dist_to_center=dist.cuda())
    g0 = torch.Generator().manual_seed(9)
    douts = [torch.randn(bs, 2, generator=g0) * 1e-3 for _ in outs]
    torch.autograd.backward(outs, [d.cuda() for d in douts])
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = O.encode(coords, enc.B.cpu(), "gauss")
    routs = O.multiscale_forward(P, x, 8, dist, bounds)
    for a, b in zip(outs, routs):
        assert rel(a, b) <= TOL
    grs = torch.autograd.grad(routs, list(P.values()), grad_outputs=douts, allow_unused=True)
    named = dict(model.named_parameters())
    frac_masked = float(((dist < 0) | (dist > 0.45)).float().mean())
    assert 0.2 < frac_masked < 0.95          # the masks actually bite in this draw
    for k, ref in zip(P.keys(), grs):
        if ref is None:
            assert named[k].grad is None, k
        else:
            assert rel(named[k].grad, ref) <= 1.5e-3, (k, rel(named[k].grad, ref))


def _gabor(inr, condition=False):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("gabor_tanh")
    if condition:
        # reference init: gamma ~ Gamma(6/9, 1) and |x|^2 + |mu|^2 ~ 430, so most envelopes underflow and only the few
        # features with gamma < 0.02 carry signal.  For the gradient checks every feature is made to matter.
        for i in range(net["network_depth"] + 1):
            sd[f"filters.{i}.gamma"] = sd[f"filters.{i}.gamma"] * 0.01 + 1e-3
            sd[f"filters.{i}.mu"] = sd[f"filters.{i}.mu"] * 0.5
    plan = inr.Plan("Gabor", net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    return plan, eng, net, loss_kind, opts, sd, encB, coords, gt


@pytest.mark.parametrize("condition", [False, True])
def test_gabor_forward_per_stage(inr, condition):
    """GaborNet (reference mfn.py:96-162): every stage's z_i and the output against the oracle, with the reference
    initialisation (sparse live features) and with a conditioned state where all envelopes are O(1)."""
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _gabor(inr, condition)
    L, bs = net["network_depth"], coords.shape[0]
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.mfn_forward(sd, x, L, True, trace=tr)
    out = eng.forward(coords.cuda(), train=True)
    for i in range(L + 1):
        assert rel(eng.read_mfn_image("z", i, bs)[:bs], tr[i]) <= 1.5e-3, f"z{i}"
    assert rel(out, out_ref) <= 1.5e-3
    # teacher-forced stage in fp64 from the engine's own z_{i-1}
    x64 = x.double()
    for i in (1, L):
        zin = eng.read_mfn_image("z", i - 1, bs)[:bs].cpu().double()
        sd64 = {k: v.double() for k, v in sd.items()}
        f = O._filter(sd64, i, x64, True)
        h = zin @ sd64[f"linear.{i-1}.weight"].t() + sd64[f"linear.{i-1}.bias"]
        assert rel(eng.read_mfn_image("z", i, bs)[:bs], f * h) <= 1e-3, i


def test_gabor_gradients_conditioned(inr):
    """Every parameter gradient of GaborNet, incl. d mu and d gamma from the q-reductions, vs torch autograd on the oracle."""
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _gabor(inr, True)
    L, bs = net["network_depth"], coords.shape[0]
    x = O.encode(coords, encB, "gauss")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.mfn_forward(P, x, L, True)
    val, dout = loss_and_grad(loss_kind, opts, o.detach(), gt, coords)
    gr = dict(zip(P.keys(), torch.autograd.grad(o, list(P.values()), grad_outputs=dout)))
    eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
    assert abs(float(eng.loss_out) - float(val)) <= 1e-4 * float(val)
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for k in sd:
        assert rel(gv[k], gr[k]) <= 2e-3, (k, rel(gv[k], gr[k]))


def test_gabor_fused_steps_vs_reference_golden(inr):
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _gabor(inr, False)
    bs = coords.shape[0]
    gold = G.load_golden("gabor_tanh")
    for step in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
        assert abs(float(eng.loss_out) - gold["losses"][step]) <= 2e-4 * gold["losses"][step], step
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        dg = G.tensor_digest(eng.params[off:off + rows * cols].cpu())
        assert abs(dg["l2"] - gold["final"][k]["l2"]) <= 1e-4 * gold["final"][k]["l2"], k


def test_kgabor_module_ignores_dist_like_the_reference(inr):
    from mri_implicit_neural_representations_b200.modules import GaborNet, KGaborNet
    torch.manual_seed(3)
    a = KGaborNet(dict(G.NET_MFN)).to("cuda")
    keys = list(a.state_dict().keys())
    assert keys[-4:] == ["filters.8.mu", "filters.8.gamma", "filters.8.linear.weight", "filters.8.linear.bias"]
    x = torch.rand(300, G.NET_MFN["network_input_size"], device="cuda") * 2 - 1
    d = torch.rand(300, device="cuda")
    with torch.no_grad():
        assert torch.equal(a(x, d), a(x, None))
    torch.manual_seed(3)
    b = GaborNet(dict(G.NET_MFN)).to("cuda")
    with torch.no_grad():
        assert torch.equal(a(x, d), b(x))


def test_gabor_full_size_per_coil_tanh_tv_config3(inr):
    """BASELINE configs[2] size: GaborNet on one whole coil of a 320 x 320 k-space slice per batch (bs 102 400 = 800 row
    tiles), tanh loss on the sampled rows of a grid-2*1 mask plus TV over all rows (reference src/train.py:172-182).
    (a) output, loss and every gradient (d mu / d gamma included) against the oracle on the whole batch -- the TV term is
    sign-based, so it is taken teacher-forced on the engine's own output; (b) size-independent properties: the same step
    twice gives the same bits, and flipping the image upside down (blocks of W rows reversed) leaves every row's output
    bit-identical and loss / gradients unchanged up to the order of the fixed-order reductions."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, _, _, _ = case_setup("gabor_tanh")
    for i in range(net["network_depth"] + 1):                    # conditioned state: every envelope is O(1), see _gabor
        sd[f"filters.{i}.gamma"] = sd[f"filters.{i}.gamma"] * 0.01 + 1e-3
        sd[f"filters.{i}.mu"] = sd[f"filters.{i}.mu"] * 0.5
    H = W = 320
    bs, L, weight = H * W, net["network_depth"], 0.5
    g = torch.Generator().manual_seed(12)
    ky, kx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    coords = torch.stack([torch.full((bs,), 0.2), kx.reshape(-1), ky.reshape(-1)], 1)
    gt = torch.rand(bs, 2, generator=g) * 0.8 + 0.1
    mask = ((torch.arange(bs) // W) % 2 == 0)
    lopts = dict(opts or {})
    lopts["tv"] = (H, W, weight)

    def grad_step(c, y, m):
        eng = inr.ChainEngine(inr.Plan("Gabor", net, enc_cfg), max_batch=bs, lr=G.LR)
        eng.load_tensors(list(sd.values()))
        eng.set_encoder(encB)
        out = torch.empty(bs, 2, device="cuda")
        eng.grad_step(loss_kind, c.cuda(), y.cuda(), bs, mask=m.to(torch.uint8).cuda(), loss_opts=lopts, out=out)
        torch.cuda.synchronize()
        return eng, out.cpu(), eng.grads.clone(), float(eng.loss_out)

    eng, out, grads, loss = grad_step(coords, gt, mask)
    x = O.encode(coords, encB, "gauss")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.mfn_forward(P, x, L, True)
    assert rel(out, o.detach()) <= 1.5e-3
    val_tv, g_tv = O.loss_tv(out, H, W, weight)
    val, dsel = O.LOSS_TRAIN["tanh"](o.detach()[mask], gt[mask])
    dout = g_tv.clone()
    dout[mask] += dsel
    assert float(g_tv.norm()) > 0.05 * float(dsel.norm())          # the TV term matters in this test
    assert abs(loss - float(val + val_tv)) <= 1e-3 * abs(float(val + val_tv)), (loss, float(val), float(val_tv))
    gr = dict(zip(P.keys(), torch.autograd.grad(o, list(P.values()), grad_outputs=dout)))
    gv = dict(zip(sd.keys(), eng._views(grads)))
    for k in sd:
        assert rel(gv[k], gr[k]) <= 2e-3, (k, rel(gv[k], gr[k]))
    _, out2, grads2, loss2 = grad_step(coords, gt, mask)
    assert torch.equal(out2, out) and torch.equal(grads2, grads) and loss2 == loss
    flip = torch.arange(bs).view(H, W).flip(0).reshape(-1)
    mask_f = mask[flip]
    _, out_f, grads_f, loss_f = grad_step(coords[flip], gt[flip], mask_f)
    assert torch.equal(out_f, out[flip])
    assert abs(loss_f - loss) <= 1e-5 * abs(loss)
    assert rel(grads_f, grads) <= 1e-4
