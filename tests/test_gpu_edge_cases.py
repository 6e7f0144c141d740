"""Edge cases through the C ABI: single-row, sub-tile and ragged batches for every model family (the engine pads to
128-row tiles internally), an all-false row mask, and loud failures for unsupported requests."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


@pytest.mark.parametrize("case,tol", [("siren_l2", 1e-3), ("ffn_l2", 1e-3), ("fourier_l2", 1.5e-3), ("gabor_tanh", 1.5e-3)])
@pytest.mark.parametrize("bs", [1, 127, 129])
def test_tiny_and_ragged_batches_forward_and_step(inr, case, tol, bs):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(case)
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=256, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    c, y = coords[:bs], gt[:bs]
    out = eng.forward(c.cuda(), train=False)
    ref = O.model_forward(model_kind, sd, O.encode(c, encB, enc_cfg["embedding"]), net)
    assert out.shape == (bs, 2)
    assert rel(out, ref) <= tol
    eng.train_step(loss_kind, c.cuda(), y.cuda(), bs, loss_opts=opts)
    torch.cuda.synchronize()
    val, _ = O.LOSS_TRAIN[loss_kind](ref, y)
    assert abs(float(eng.loss_out) - float(val)) <= 2e-3 * abs(float(val)) + 1e-9
    assert torch.isfinite(eng.params).all()


@pytest.mark.parametrize("case", ["wire_l2", "wire2d_l2"])
def test_complex_models_sub_tile_batch(inr, case):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(case)
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=128, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    bs = 5
    sd64 = {k: (v.to(torch.complex128) if v.is_complex() else v.double()) for k, v in sd.items()}
    out = eng.forward(coords[:bs].cuda(), train=False)
    o32 = O.model_forward(model_kind, sd, coords[:bs], net)
    o64 = O.model_forward(model_kind, sd64, coords[:bs].double(), net)
    assert rel(out, o64) <= 4 * rel(o32, o64) + 2e-4
    eng.train_step(loss_kind, coords[:bs].cuda(), gt[:bs].cuda(), bs, loss_opts=opts)
    torch.cuda.synchronize()
    assert torch.isfinite(eng.params).all() and torch.isfinite(eng.loss_out).all()


def test_all_false_mask_gives_zero_gradient_and_no_nan_in_parameters(inr):
    """No row enters the loss (the reference would take the mean of an empty tensor -> NaN loss); the engine reports a
    zero masked count, produces zero gradients and leaves the parameters finite."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, _ = case_setup("siren_l2")
    plan = inr.Plan(model_kind, net, enc_cfg)
    bs = 300
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    mask = torch.zeros(bs, dtype=torch.uint8, device="cuda")
    g = eng.grad_step("L2", coords[:bs].cuda(), gt[:bs].cuda(), bs, mask=mask)
    torch.cuda.synchronize()
    assert float(g.abs().max()) == 0.0
    assert float(eng.scalars(bs)[4]) == 0.0          # SC_COUNT


def test_unsupported_requests_fail_loudly(inr):
    net = dict(G.NET_256)
    with pytest.raises(Exception):
        inr.Plan("SIREN", dict(net, network_last_linear=False), G.ENC_GAUSS)          # sine output layer: not built
    with pytest.raises(Exception):
        inr.Plan("SIREN", net, {"embedding": "LogF", "scale": 4, "embedding_size": 256, "coordinates_size": 3})
    with pytest.raises(Exception):
        inr.Plan("WIRE2D", dict(G.NET_W2D, last_tanh=True), G.ENC_NONE)
    with pytest.raises(NotImplementedError):
        inr.Plan("NoSuchModel", net, G.ENC_GAUSS)
    plan = inr.Plan("SIREN", net, G.ENC_GAUSS)
    eng = inr.ChainEngine(plan, max_batch=128)
    with pytest.raises(Exception):                                                      # TV needs bs == H * W and `out`
        eng.train_step("L2", torch.zeros(100, 3, device="cuda"), torch.zeros(100, 2, device="cuda"), 100,
                       loss_opts={"tv": (8, 8)})
