"""GPU parity tests of the LogF positional encoder (reference src/models/networks.py:14-16,24-29) fused into the SIREN / FFN
step through the C ABI: the encoder image itself, forward per layer, every gradient (the first-layer weight has 6 n = 252
columns in the reference's layout while the GEMM runs on K = 256), fused steps against the golden run of the reference."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup, loss_and_grad

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _engine(inr):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("siren_logf")
    plan = inr.Plan(model_kind, net, enc_cfg)
    assert plan.wide and plan.tensors[0][2] == 252                 # reference shape of the first-layer weight: [256, 252]
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    return plan, eng, net, loss_kind, opts, sd, encB, coords, gt


def test_logf_forward_per_layer_and_gradients(inr):
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _engine(inr)
    n, depth = coords.shape[0], net["network_depth"]
    x = O.encode(coords, encB, "LogF")
    assert x.shape[1] == 252
    tr = []
    out_ref = O.siren_forward(sd, x, depth, trace=tr)
    out = eng.forward(coords.cuda(), train=True)
    # the encoder image: 252 features in the reference's order, then four zero columns (fp16 image of sin / cos values)
    ximg = eng.read_mfn_image("x", 0, n)[:n]
    assert ximg.shape[1] == 256
    assert float(ximg[:, 252:].abs().max()) == 0.0
    assert float((ximg[:, :252].cpu().double() - x.double()).abs().max()) <= 1e-3      # fp16 rounding of values in [-1, 1]
    for i in range(depth - 1):
        assert rel(eng.read_mfn_image("z", i, n)[:n], tr[i][1]) <= 1e-3, f"layer {i}"
    assert rel(out, out_ref) <= 1e-3
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.siren_forward(P, x, depth)
    val, dout = loss_and_grad(loss_kind, opts, o.detach(), gt, coords)
    gr = dict(zip(P.keys(), torch.autograd.grad(o, list(P.values()), grad_outputs=dout)))
    for _ in range(2):                       # the first pass calibrates the per-layer gradient scales
        eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), n)
    assert abs(float(eng.loss_out) - float(val)) <= 1e-3 * float(val)
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    assert tuple(gv["model.0.linear.weight"].shape) == (256, 252)
    for k in sd:
        assert rel(gv[k], gr[k]) <= 2e-3, (k, rel(gv[k], gr[k]))


def test_logf_fused_steps_match_the_reference_golden(inr):
    plan, eng, net, loss_kind, opts, sd, encB, coords, gt = _engine(inr)
    n = coords.shape[0]
    gold = G.load_golden("siren_logf")
    losses = []
    for _ in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, coords.cuda(), gt.cuda(), n)
        losses.append(float(eng.loss_out))
    assert abs(losses[0] - gold["losses"][0]) <= 1e-3 * gold["losses"][0], (losses, gold["losses"])
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 6e-3 * b, (losses, gold["losses"])
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        got = eng.params[off:off + rows * cols].cpu().double()
        assert abs(float(got.norm()) - gold["final"][k]["l2"]) <= 1e-3 * gold["final"][k]["l2"], k


def test_logf_through_the_drop_in_trainer_and_unsupported_uses(inr):
    """FusedTrainer with the module + Positional_Encoder of the drop-in; LogF with an MFN fails loudly."""
    from mri_implicit_neural_representations_b200.modules import SIREN, Positional_Encoder
    from mri_implicit_neural_representations_b200.trainer import FusedAdam, FusedTrainer
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("siren_logf")
    torch.manual_seed(1)
    enc = Positional_Encoder(enc_cfg, "cuda")
    assert torch.equal(enc.B.cpu(), encB)
    m = SIREN(dict(net)).to("cuda")
    m.load_state_dict(sd)
    optim = FusedAdam(m, lr=G.LR)
    tr = FusedTrainer(m, enc, optim, "L2", coords.shape[0], coords.cuda(), gt.cuda(), use_graph=False)
    l0 = float(tr.step())
    gold = G.load_golden("siren_logf")
    assert abs(l0 - gold["losses"][0]) <= 1e-3 * gold["losses"][0]
    with pytest.raises(Exception):
        inr.Plan("Fourier", dict(G.NET_MFN, network_input_size=252), enc_cfg)
    with pytest.raises(Exception):
        inr.Plan("SIREN", dict(net, network_input_size=256), enc_cfg)          # 6 * int(256 / 6) = 252, not 256
