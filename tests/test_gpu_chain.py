"""GPU parity tests of the fused chain kernels (SIREN / FFN) through the C ABI.

Oracle: oracle/inr_oracle.py (CPU, fp32) on the same seeded inputs; golden digests from the
reference itself (tests/golden).  Tolerance: north_star's <= 1e-3 relative L2 per layer."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup, loss_and_grad

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


@pytest.fixture(autouse=True, params=["rows128", "transposed"])
def tiling(request, monkeypatch):
    """Every test of this file runs on both tilings of the width-256 chain: 128-row tiles (chain_fwd.cu / chain_bwd.cu) and
    the transposed short tiles the library picks below one wave of rows (chain_t.cu; these cases: 16-row tiles)."""
    monkeypatch.setenv("INR_CHAIN_T", "0" if request.param == "rows128" else "1")
    return request.param


def _engine(inr, name, bs=None):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs or coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    return plan, eng


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_umma_selftest(inr, mode):
    err, ref = inr.selftest_umma(mode, 0)
    assert err <= 2e-3 * max(ref, 1.0), (mode, err, ref)


@pytest.mark.parametrize("name", ["siren_l2", "siren_tanh", "ffn_l2"])
def test_forward_per_layer(inr, name):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan, eng = _engine(inr, name)
    bs, depth = coords.shape[0], net["network_depth"]
    x = O.encode(coords, encB, enc_cfg["embedding"])
    tr = []
    out_ref = O.model_forward(model_kind, sd, x, net, trace=tr)
    out = eng.forward(coords.cuda(), train=True)
    assert rel(eng.read_image("h", 0, bs)[:bs], x) <= TOL
    for l in range(depth - 1):
        assert rel(eng.read_image("h", l + 1, bs)[:bs], tr[l][1]) <= TOL, f"H{l+1}"
    assert rel(out, out_ref) <= TOL


@pytest.mark.parametrize("name", ["siren_l2", "ffn_l2"])
def test_forward_dense_input_and_inference(inr, name):
    """encoder 'none' plan: the [bs,512] embedding is passed in as the reference's model(x) receives it."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan = inr.Plan(model_kind, net, {"embedding": "none"})
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0])
    eng.load_tensors(list(sd.values()))
    x = O.encode(coords, encB, "gauss")
    out_ref = O.model_forward(model_kind, sd, x, net)
    out = eng.forward(x.cuda(), train=False)
    assert rel(out, out_ref) <= TOL


@pytest.mark.parametrize("name", ["siren_l2", "siren_tanh", "ffn_l2"])
def test_backward_external_dout(inr, name):
    """inr_backward (autograd path): dgrad images and every weight / bias gradient vs the explicit oracle."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan, eng = _engine(inr, name)
    bs, depth = coords.shape[0], net["network_depth"]
    x = O.encode(coords, encB, enc_cfg["embedding"])
    tr = []
    out_ref = O.model_forward(model_kind, sd, x, net, trace=tr)
    _, dout = loss_and_grad(loss_kind, opts, out_ref, gt, coords)
    eng.forward(coords.cuda(), train=True)
    if model_kind == "SIREN":
        grads_ref, dzs = O.siren_backward(sd, x, tr, dout, depth, net.get("last_tanh", False))
    else:
        # ReLU': teacher-forced on the engine's own masks; they may differ from the oracle's only where
        # |z| is within fp16-operand rounding of zero (asserted rare)
        masks = [eng.read_image("d", l, bs)[:bs].cpu() for l in range(depth - 1)]
        for l in range(depth - 1):
            flips = (masks[l] != (tr[l][0] > 0).float()).float().mean()
            assert float(flips) < 5e-3, (l, float(flips))
        grads_ref, dzs = O.ffn_backward(sd, x, tr, dout, depth, masks=masks)
    g = eng.backward(dzs[depth - 1].cuda())          # contract: dL/dz_last
    S = float(eng.scalars(bs)[1])
    for l in range(depth - 1):
        assert rel(eng.read_image("dz", l, bs)[:bs] / S, dzs[l]) <= TOL, f"dZ{l}"
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = g[off:off + rows * cols].view(grads_ref[k].shape)
        assert rel(gv, grads_ref[k]) <= TOL, k


@pytest.mark.parametrize("name", ["siren_l2", "siren_tanh", "siren_l1", "ffn_l2", "ffn_msle"])
def test_fused_train_steps_vs_reference_golden(inr, name):
    """Three fused steps (forward + loss + backward + Adam) against the reference's own run
    (tests/golden/<name>.json): loss of every step and final parameters."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan, eng = _engine(inr, name)
    bs = coords.shape[0]
    gold = G.load_golden(name)
    cd, gd = coords.cuda(), gt.cuda()
    out = torch.empty(bs, 2, device="cuda")
    for step in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, cd, gd, bs, loss_opts=opts, out=out if step == 0 else None)
        loss = float(eng.loss_out)
        # step 0 is a pure forward+loss check; later steps inherit Adam's sign-like first updates
        # (m/sqrt(v) ~ +-1), which amplify fp16-operand gradient noise on near-zero entries
        tol = 5e-4 if step == 0 else 6e-3
        assert abs(loss - gold["losses"][step]) <= tol * abs(gold["losses"][step]), (step, loss)
        if step == 0:
            # sampled elements: within 3e-3 of the output rms (the 1e-3 bar is a relative L2 over the tensor; single
            # elements of an fp16-operand SIREN scatter a few times wider)
            assert not G.digest_close(gold["out"], G.tensor_digest(out.cpu()), 2e-3, atol_scale=1.5)
    assert int(eng.step) == G.N_ADAM_STEPS
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        dg = G.tensor_digest(eng.params[off:off + rows * cols].cpu())
        gf = gold["final"][k]
        assert abs(dg["l2"] - gf["l2"]) <= 1e-3 * gf["l2"], k


def test_ragged_batch_and_mask(inr):
    """bs not a multiple of 128 and a row mask (src/train.py:172-177): loss and gradients vs oracle."""
    name = "siren_l2"
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, _ = case_setup(name)
    bs = 333
    coords, gt = coords[:bs], gt[:bs]
    mask = (torch.arange(bs) % 3 != 0)
    plan, eng = _engine(inr, name, bs=bs)
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.model_forward(model_kind, sd, x, net, trace=tr)
    val, dsel = O.loss_l2(out_ref[mask], gt[mask])
    dout = torch.zeros_like(out_ref)
    dout[mask] = dsel
    grads_ref, _ = O.siren_backward(sd, x, tr, dout, net["network_depth"])
    eng.hyper[0:1].fill_(0.0)     # lr 0: parameters stay, gradients are still produced
    eng.train_step("L2", coords.cuda(), gt.cuda(), bs, mask=mask.to(torch.uint8).cuda())
    assert abs(float(eng.loss_out) - float(val)) <= 1e-3 * float(val)
    # first Adam moment after one step = (1-beta1) * g
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = eng.exp_avg[off:off + rows * cols].view(grads_ref[k].shape) / 0.1
        assert rel(gv, grads_ref[k]) <= TOL, k


def test_siren_full_size_batch_config0(inr, tiling):
    """BASELINE configs[0] size (SIREN d4 w256, gauss encoder, L2, batch 10 000 = 79 row tiles / 125 transposed tiles, last
    one ragged) through the fused step (reference src/train.py:158-192).  (a) every output row, the loss and every
    gradient against the oracle on the whole batch, (b) size-independent properties: a row permutation of the batch leaves
    loss and gradients unchanged up to the order of the fixed-order reductions, and two engines fed the same three steps
    end bit-identical (a tile read before it was complete would show up here)."""
    name = "siren_l2"
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, _, _, _ = case_setup(name)
    bs = 10000
    g = torch.Generator().manual_seed(17)
    coords = torch.rand(bs, 3, generator=g) * 2 - 1
    gt = torch.randn(bs, 2, generator=g) * 0.1
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.model_forward(model_kind, sd, x, net, trace=tr)
    val, dout = O.loss_l2(out_ref, gt)
    grads_ref, _ = O.siren_backward(sd, x, tr, dout, net["network_depth"])

    def one_step(c, y):
        plan, eng = _engine(inr, name, bs=bs)
        eng.hyper[0:1].fill_(0.0)                                  # lr 0: the first Adam moment is (1 - beta1) * g
        out = torch.empty(bs, 2, device="cuda")
        eng.train_step("L2", c.cuda(), y.cuda(), bs, out=out)
        torch.cuda.synchronize()
        return plan, eng, out.cpu()

    plan, eng, out = one_step(coords, gt)
    assert rel(out, out_ref) <= TOL
    assert abs(float(eng.loss_out) - float(val)) <= 1e-3 * float(val)
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = eng.exp_avg[off:off + rows * cols].view(grads_ref[k].shape) / 0.1
        assert rel(gv, grads_ref[k]) <= TOL, k
    perm = torch.randperm(bs, generator=g)
    _, eng_p, out_p = one_step(coords[perm], gt[perm])
    assert torch.equal(out_p, out[perm])                           # a row's output does not depend on its tile or lane
    assert abs(float(eng_p.loss_out) - float(eng.loss_out)) <= 1e-5 * abs(float(eng.loss_out))
    assert rel(eng_p.exp_avg, eng.exp_avg) <= 1e-4
    engs = [_engine(inr, name, bs=bs)[1] for _ in range(2)]
    cd, yd = coords.cuda(), gt.cuda()
    for e in engs:
        for _ in range(3):
            e.train_step("L2", cd, yd, bs)
    torch.cuda.synchronize()
    assert torch.equal(engs[0].params, engs[1].params)
    assert torch.isfinite(engs[0].params).all()
