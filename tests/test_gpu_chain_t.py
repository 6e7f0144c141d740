"""GPU parity tests of the transposed small-batch chain kernels (csrc/chain_t.cu) at every tile height they use.

tests/test_gpu_chain.py checks them against the oracle and the reference's golden runs at the oracle's sizes (16-row
tiles).  Here the batch is large enough for 32 .. 80-row tiles (up to BASELINE configs[0]'s 10 000 rows): the checker is
the oracle on a row sample of the forward, and the 128-row-tile kernels -- themselves oracle-checked -- for the whole
step: same fp16 operands, so the two tilings may differ by fp32 summation order only."""
import pytest
import torch

from oracle import inr_oracle as O
from oracle.cases import case_setup

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _data(bs, seed, k0=None):
    g = torch.Generator().manual_seed(seed)
    coords = torch.rand(bs, 3, generator=g) * 2 - 1
    gt = torch.rand(bs, 2, generator=g) * 0.8 + 0.1
    return coords, gt


def _engine(inr, name, bs, enc=None):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, *_ = case_setup(name)
    plan = inr.Plan(model_kind, net, enc or enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=2e-5)
    eng.load_tensors(list(sd.values()))
    if (enc or enc_cfg)["embedding"] == "gauss":
        eng.set_encoder(encB)
    return plan, eng, (model_kind, net, sd, encB, loss_kind, opts)


# rows -> tile height on 148 SMs: 16 * ceil(bs / (148 * 16))
SIZES = [(2369, 32), (5000, 48), (7500, 64), (9999, 80), (10000, 80), (11840, 80)]


@pytest.mark.parametrize("bs,nr", SIZES)
@pytest.mark.parametrize("name", ["siren_l2", "ffn_l2"])
def test_tile_height_and_forward_vs_oracle(inr, monkeypatch, name, bs, nr):
    monkeypatch.setenv("INR_CHAIN_T", "1")
    plan, eng, (kind, net, sd, encB, *_rest) = _engine(inr, name, bs)
    lay = plan.workspace_layout(bs)
    if torch.cuda.get_device_properties(0).multi_processor_count == 148:
        assert lay["tile_rows"] == nr and lay["lb"] == nr * 16 + 16 and lay["n_tiles"] == -(-bs // nr)
    assert plan.workspace_layout(148 * 80 + 1)["tile_rows"] == 128            # a full wave of rows: 128-row tiles
    coords, _ = _data(bs, 5)
    out = eng.forward(coords.cuda(), train=True)
    idx = torch.cat([torch.arange(0, 200), torch.arange(bs - 200, bs), torch.randperm(bs)[:400]])
    x = O.encode(coords[idx], encB, "gauss")
    tr = []
    ref = O.model_forward(kind, sd, x, net, trace=tr)
    assert rel(out[idx.cuda()], ref) <= 1e-3
    assert rel(eng.read_image("h", 0, bs)[idx.cuda()], x) <= 1e-3
    for l in range(net["network_depth"] - 1):
        assert rel(eng.read_image("h", l + 1, bs)[idx.cuda()], tr[l][1]) <= 1e-3, l


@pytest.mark.parametrize("bs,nr", SIZES)
@pytest.mark.parametrize("name,loss", [("siren_l2", "L2"), ("siren_tanh", "tanh"), ("ffn_l2", "L1"), ("ffn_msle", "MSLE")])
def test_grad_step_matches_row_tile_kernels(inr, monkeypatch, name, loss, bs, nr):
    """loss, outputs, every saved image and the full gradient of one fused step, transposed tiles vs 128-row tiles."""
    coords, gt = _data(bs, 7)
    mask = (torch.arange(bs) % 5 != 0).to(torch.uint8)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("INR_CHAIN_T", mode)
        plan, eng, _ = _engine(inr, name, bs)
        out = torch.empty(bs, 2, device="cuda")
        g = eng.grad_step(loss, coords.cuda(), gt.cuda(), bs, mask=mask.cuda(), out=out).clone()
        depth = plan.desc.depth
        imgs = {("h", l): eng.read_image("h", l, bs)[:bs].cpu() for l in range(depth)}
        imgs.update({("d", l): eng.read_image("d", l, bs)[:bs].cpu() for l in range(depth - 1)})
        imgs.update({("dz", l): eng.read_image("dz", l, bs)[:bs].cpu() for l in range(depth - 1)})
        res[mode] = (float(eng.loss_out), out.cpu(), g.cpu(), imgs, plan)
    (l0, o0, g0, i0, plan), (l1, o1, g1, i1, _) = res["0"], res["1"]
    # ReLU chain: identical fp16 operands, the tilings differ by fp32 summation order and a rare fp16 rounding flip.
    # SIREN: a one-ulp flip of a stored fp16 activation moves the next layer's phase w0 * (W h) by ~1e-4, so the two tilings
    # decorrelate to the level of the fp16-operand error itself (each is within 1e-3 of the oracle, see the test above).
    siren = name.startswith("siren")
    m = {"loss": abs(l0 - l1) / abs(l0), "out": rel(o1, o0), "grad": rel(g1, g0)}
    m.update({k: rel(i1[k], i0[k]) for k in i0})
    m.update({("g", layer, is_bias): rel(g1[off:off + rows * cols], g0[off:off + rows * cols])
              for off, rows, cols, layer, is_bias in plan.tensors})
    print({k: f"{v:.2e}" for k, v in m.items()})
    tol = {"loss": 2e-4 if siren else 1e-5, "out": 1e-3 if siren else 1e-5, "grad": 3e-3 if siren else 1e-3}
    for k, v in m.items():
        bound = tol.get(k, {"h": 1e-3 if siren else 5e-4, "d": 2e-3 if siren else 5e-4, "dz": 3e-3 if siren else 2e-3,
                            "g": 5e-3 if siren else 2e-3}.get(k[0] if isinstance(k, tuple) else k))
        assert v <= bound, (k, v, bound)


@pytest.mark.parametrize("bs", [5000, 10000])
def test_train_steps_track_row_tile_kernels(inr, monkeypatch, bs):
    """Twenty fused steps (loss + backward + Adam, device cursor, sample-mode batches out of a resident set)."""
    n = 3 * bs + 123
    coords, gt = _data(n, 9)
    losses = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("INR_CHAIN_T", mode)
        plan, eng, _ = _engine(inr, "siren_l2", bs)
        cd, gd = coords.cuda(), gt.cuda()
        ls = []
        for step in range(20):
            eng.cursor.fill_((step % 3) * bs)
            eng.train_step("L2", cd, gd, bs, use_cursor=True)
            ls.append(float(eng.loss_out))
        losses[mode] = (ls, eng.params.clone().cpu())
    for a, b in zip(losses["0"][0], losses["1"][0]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses["0"][0], losses["1"][0])
    assert losses["1"][0][-1] < losses["1"][0][0]
    assert rel(losses["1"][1], losses["0"][1]) <= 1e-4


def test_dense_input_tv_and_external_dout(inr, monkeypatch):
    """encoder 'none' plans, the per-coil TV term (bs = H * W) and the autograd face's backward on the short tiles."""
    H, W = 60, 100
    bs = H * W
    coords, gt = _data(bs, 11)
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, *_ = case_setup("siren_l2")
    x = O.encode(coords, encB, "gauss").cuda()
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("INR_CHAIN_T", mode)
        plan, eng, _ = _engine(inr, "siren_l2", bs, enc={"embedding": "none"})
        out = torch.empty(bs, 2, device="cuda")
        g = eng.grad_step("L2", None, gt.cuda(), bs, x=x, out=out, loss_opts={"tv": (H, W, 1e-2)}).clone().cpu()
        l_tv = float(eng.loss_out)
        y = eng.forward(x, train=True)
        gb = eng.backward(torch.ones(bs, 2, device="cuda") * 1e-3).clone().cpu()
        res[mode] = (l_tv, out.cpu(), g, y.cpu(), gb)
    a, b = res["0"], res["1"]
    m = [abs(a[0] - b[0]) / abs(a[0]), rel(b[1], a[1]), rel(b[3], a[3]), rel(b[2], a[2]), rel(b[4], a[4])]
    print(["%.2e" % v for v in m])
    assert m[0] <= 2e-4 and m[1] <= 1e-3 and m[2] <= 1e-3        # SIREN: see test_grad_step_matches_row_tile_kernels
    assert m[3] <= 3e-3 and m[4] <= 3e-3
