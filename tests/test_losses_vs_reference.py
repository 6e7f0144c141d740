"""src/metrics/losses.py (the drop-in loss classes the entry points construct) against the UNMODIFIED reference classes
(src/metrics/losses.py there): values and input gradients on the same random inputs.  Live test."""
import importlib.util
import os

import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = {"hdr_ff_sigma": 1.5, "hdr_eps": 1e-2, "hdr_ff_factor": 0.3, "min_sample": 40}


@pytest.fixture(scope="module")
def mods():
    ref = ref_shims.load("metrics.losses")
    spec = importlib.util.spec_from_file_location("inr_src_losses", os.path.join(ROOT, "src", "metrics", "losses.py"))
    ours = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ours)
    return ref, ours


def _inputs(n=600, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, 2, generator=g) * 0.3).requires_grad_(True)
    y = torch.randn(n, 2, generator=g) * 0.3
    k = torch.rand(n, 3, generator=g) * 2 - 1
    return x, y, k


def _value_and_grad(fn, x):
    out = fn(x)
    val = out[0] if isinstance(out, tuple) else out
    (g,) = torch.autograd.grad(val.sum(), x)
    return float(val.sum()), g


def _check(ref_fn, our_fn, tol=1e-5):
    x, *_ = _inputs()
    vr, gr = _value_and_grad(ref_fn, x)
    vo, go = _value_and_grad(our_fn, x)
    assert abs(vr - vo) <= tol * max(abs(vr), 1e-12), (vr, vo)
    assert float((gr - go).norm()) <= tol * max(float(gr.norm()), 1e-12)


def test_pointwise_losses(mods):
    R, O = mods
    _, y, k = _inputs()
    _check(lambda x: R.MSLELoss()(x.abs(), y.abs()), lambda x: O.MSLELoss()(x.abs(), y.abs()))
    _check(lambda x: R.TanhL2Loss()(x, y, k), lambda x: O.TanhL2Loss()(x, y, k))
    _check(lambda x: R.LogSpaceLoss(OPTS)(x, y), lambda x: O.LogSpaceLoss(OPTS)(x, y))
    _check(lambda x: R.TLoss()(x, y), lambda x: O.TLoss()(x, y))
    _check(lambda x: R.FocalFrequencyLoss()(x, y), lambda x: O.FocalFrequencyLoss()(x, y))


def test_hdr_separable_form_equals_reference(mods):
    R, O = mods
    _, y, k = _inputs(n=200)
    x = _inputs(n=200)[0]
    vr, gr = _value_and_grad(lambda t: R.HDRLoss_FF(OPTS)(t, y, k), x)
    vo, go = _value_and_grad(lambda t: O.HDRLoss_FF(OPTS)(t, y, k), x)
    assert abs(vr - vo) <= 2e-5 * abs(vr) and float((gr - go).norm()) <= 2e-5 * float(gr.norm())


def test_center_loss_with_the_same_rng_state(mods):
    R, O = mods
    _, y, k = _inputs()
    x = _inputs()[0]
    torch.manual_seed(3)
    vr, gr = _value_and_grad(lambda t: R.CenterLoss(OPTS)(t, y, k), x)
    torch.manual_seed(3)
    vo, go = _value_and_grad(lambda t: O.CenterLoss(OPTS)(t, y, k), x)
    assert abs(vr - vo) <= 1e-5 * abs(vr), (vr, vo)
    assert float((gr - go).norm()) <= 1e-5 * float(gr.norm())


def test_consistency_and_tv(mods):
    R, O = mods
    g = torch.Generator().manual_seed(5)
    outs = [torch.randn(500, 2, generator=g).requires_grad_(True) for _ in range(4)]
    dist = torch.rand(500, generator=g) * 1.4
    bounds = [(0, 0.3), (0, 0.6), (0, 0.9), (0, 5)]
    lr = R.ConsistencyLoss(bounds)(outs, dist)
    lo = O.ConsistencyLoss(bounds)(outs, dist)
    assert abs(float(lr) - float(lo)) <= 1e-6 * abs(float(lr))
    for a, b in zip(torch.autograd.grad(lr, outs[1:]), torch.autograd.grad(lo, outs[1:])):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-9)
    img = torch.randn(16, 20, 2, generator=g)
    assert abs(float(R.tv_loss(img)) - float(O.tv_loss(img))) <= 1e-9
