"""Out-of-bounds and race screening without compute-sanitizer (closed on this pool: tools/sanitize.sh exits 86).

* guard bands: the engine's workspace, packed-operand and gradient buffers are re-seated inside larger allocations filled with
  a canary pattern; after fused steps of every kernel family (WIRE CTA-pair chains incl. an odd tile count, SIREN chain,
  wide SIREN chain, fused multi-scale step) every canary byte on both sides must be intact;
* determinism: two engines started from the same state must end bit-identical (parameters, packed operands) -- the chained
  kernels hand tiles over through acquire / release counters and cluster barriers, a race shows up as run-to-run noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu
ENC = {"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3}
GUARD = 1 << 20
PAIRS = [(0.0, 0.4), (0.0, 0.8), (0.0, 1.1), (0.0, 5.0)]
CASES = {
    "wire_hdr_masked": ("WIRE", {"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256,
                                 "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}, {"embedding": "none"}, "HDR", 2560,
                        {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}, True, False),
    "wire_l2_odd_tiles": ("WIRE", {"network_input_size": 3, "network_output_size": 2, "network_depth": 2, "network_width": 256,
                                   "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}, {"embedding": "none"}, "L2", 1100, None, False, False),
    "siren_l2": ("SIREN", {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256}, ENC, "L2", 1300,
                 None, False, False),
    "siren_w512_tanh": ("SIREN", {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 512,
                                  "last_tanh": True}, ENC, "tanh", 700, None, False, False),
    "bounded_fourier_lsl": ("BoundedFourier", {"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 256,
                                               "boundaries": [p for p in PAIRS for _ in (0, 1)]}, ENC, "LSL", 600,
                            {"hdr_eps": 1e-2, "consistency": (PAIRS, 0.1)}, False, True),
}


def _guarded(t):
    """A tensor of t's shape / dtype seated in the middle of a canary-filled allocation; returns (view, whole, offset)."""
    nbytes = t.numel() * t.element_size()
    pad = (nbytes + 1023) // 1024 * 1024
    whole = torch.full((GUARD + pad + GUARD,), 0xA5, dtype=torch.uint8, device=t.device)
    view = whole[GUARD:GUARD + nbytes].view(t.dtype).view(t.shape)
    view.copy_(t)
    return view, whole, nbytes


def _build(name):
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200 import init as pinit
    from oracle import inr_oracle as O
    model, net, enc, loss, bs, opts, masked, dist = CASES[name]
    torch.manual_seed(1)
    plan = inr.Plan(model, net, enc)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=5e-4)
    if model == "WIRE":
        tensors = [t for _, t in pinit.wire_tensors(net)]
    elif model == "BoundedFourier":
        tensors = list(O.multiscale_init({k: v for k, v in net.items() if k != "boundaries"}, bounded=True).values())
    else:
        tensors = [t for _, t in pinit.chain_tensors(model, net)]
    eng.load_tensors(tensors)
    eng.set_encoder(pinit.encoder_matrix(enc))
    g = torch.Generator().manual_seed(2)
    c = (torch.rand(bs, 3, generator=g) * 2 - 1).cuda()
    y = (torch.randn(bs, 2, generator=g) * 0.05).cuda()
    m = (torch.arange(bs, device="cuda") % 2 == 0).to(torch.uint8) if masked else None
    d = torch.sqrt(c[:, 1] ** 2 + c[:, 2] ** 2) if dist else None
    return eng, (loss, c, y, bs), dict(mask=m, loss_opts=opts, dist=d)


@pytest.mark.parametrize("name", list(CASES))
def test_guard_bands_stay_intact_and_runs_are_bit_identical(name):
    finals = []
    for rep in range(2):
        eng, args, kw = _build(name)
        guards = {}
        for attr in ("workspace", "wpack", "grads"):
            view, whole, nbytes = _guarded(getattr(eng, attr))
            setattr(eng, attr, view)
            guards[attr] = (whole, nbytes)
        eng.pack()                                   # the packed operands live in the guarded buffer now
        for _ in range(3):
            eng.train_step(*args, **kw)
        torch.cuda.synchronize()
        for attr, (whole, nbytes) in guards.items():
            assert bool((whole[:GUARD] == 0xA5).all()), f"{name}: write below {attr}"
            pad_end = whole[GUARD + nbytes:]
            assert bool((pad_end == 0xA5).all()), f"{name}: write above {attr}"
        assert torch.isfinite(eng.params).all()
        finals.append((eng.params.clone(), eng.wpack.clone()))
    assert torch.equal(finals[0][0], finals[1][0]), f"{name}: parameters differ between two identical runs"
    assert torch.equal(finals[0][1], finals[1][1]), f"{name}: packed operands differ between two identical runs"
