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
    assert inr.Plan("SIREN", dict(net, network_last_linear=False), G.ENC_GAUSS).wide  # sine output layer: built since round 2 (wide chain)
    with pytest.raises(Exception):
        inr.Plan("SIREN", net, {"embedding": "LogF", "scale": 4, "embedding_size": 256, "coordinates_size": 3})
    inr.Plan("WIRE2D", dict(G.NET_W2D, last_tanh=True), G.ENC_NONE)      # complex tanh tail: built (tests/test_gpu_wire2d.py)
    with pytest.raises(NotImplementedError):
        inr.Plan("NoSuchModel", net, G.ENC_GAUSS)
    plan = inr.Plan("SIREN", net, G.ENC_GAUSS)
    eng = inr.ChainEngine(plan, max_batch=128)
    with pytest.raises(Exception):                                                      # TV needs bs == H * W and `out`
        eng.train_step("L2", torch.zeros(100, 3, device="cuda"), torch.zeros(100, 2, device="cuda"), 100,
                       loss_opts={"tv": (8, 8)})


def test_two_wire_fits_step_concurrently_under_an_sm_budget():
    """chained launches need all their CTAs resident, so two fits stepping at the same time on one device
    used to be able to starve each other (bounded wait -> trap).  With inr_set_sm_budget(74) each chained launch takes half
    the chip: two WIRE fits on two streams run side by side and end exactly where the same fits end when run one after the
    other (the grid size does not enter the arithmetic: every reduction has a fixed order)."""
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200 import init as pinit
    net = {"network_input_size": 3, "network_output_size": 2, "network_depth": 3, "network_width": 256, "first_omega_0": 30,
           "hidden_omega_0": 30, "scale": 15}
    bs = 6000

    def make(seed):
        torch.manual_seed(seed)
        eng = inr.ChainEngine(inr.Plan("WIRE", net, {"embedding": "none"}), max_batch=bs, lr=1e-4)
        eng.load_tensors([t for _, t in pinit.wire_tensors(net)])
        g = torch.Generator().manual_seed(seed)
        return eng, (torch.rand(bs, 3, generator=g) * 2 - 1).cuda(), (torch.randn(bs, 2, generator=g) * 0.05).cuda()

    def run(concurrent):
        fits = [make(1), make(2)]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        for eng, c, y in fits:                       # calibration pass + kernel attributes on the default stream
            eng.train_step("L2", c, y, bs)
        torch.cuda.synchronize()
        for _ in range(20):
            for (eng, c, y), st in zip(fits, streams):
                with torch.cuda.stream(st if concurrent else torch.cuda.current_stream()):
                    eng.train_step("L2", c, y, bs)
            if not concurrent:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return [eng.params.clone() for eng, _, _ in fits]

    inr.set_sm_budget(74)
    try:
        together = run(True)
        apart = run(False)
    finally:
        inr.set_sm_budget(0)
    for a, b in zip(together, apart):
        assert torch.isfinite(a).all()
        assert torch.equal(a, b)
