"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/dropin_losses.json from the UNMODIFIED reference loss classes
(src/metrics/losses.py) on seeded inputs: loss value and the L2 norm / first entries of its input gradient.

    python oracle/make_golden_dropin_losses.py        (needs /root/reference; run in the build container)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OPTS = {"hdr_ff_sigma": 1.5, "hdr_eps": 1e-2, "hdr_ff_factor": 0.3, "min_sample": 40}
BOUNDS = [(0, 0.3), (0, 0.6), (0, 0.9), (0, 5)]


def inputs(n=400, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, 2, generator=g) * 0.3).requires_grad_(True)
    y = torch.randn(n, 2, generator=g) * 0.3
    k = torch.rand(n, 3, generator=g) * 2 - 1
    return x, y, k


def evaluate(L):
    """L: a module object exposing the loss classes (the reference's or this repo's drop-in)."""
    out = {}

    def rec(name, fn):
        x, y, k = inputs()
        torch.manual_seed(3)                     # CenterLoss draws torch.randperm
        val = fn(x, y, k)
        val = val[0] if isinstance(val, tuple) else val
        (g,) = torch.autograd.grad(val.sum(), x)
        out[name] = {"value": float(val.detach().sum()), "grad_l2": float(g.norm()), "grad_head": [float(v) for v in g.reshape(-1)[:4]]}
    rec("MSLE", lambda x, y, k: L.MSLELoss()(x.abs(), y.abs()))
    rec("tanh", lambda x, y, k: L.TanhL2Loss()(x, y, k))
    rec("LogSpace", lambda x, y, k: L.LogSpaceLoss(OPTS)(x, y))
    rec("HDR", lambda x, y, k: L.HDRLoss_FF(OPTS)(x, y, k))
    rec("T", lambda x, y, k: L.TLoss()(x, y))
    rec("Center", lambda x, y, k: L.CenterLoss(OPTS)(x, y, k))
    rec("FFL", lambda x, y, k: L.FocalFrequencyLoss()(x, y))
    g = torch.Generator().manual_seed(5)
    outs = [torch.randn(300, 2, generator=g).requires_grad_(True) for _ in range(4)]
    dist = torch.rand(300, generator=g) * 1.4
    val = L.ConsistencyLoss(BOUNDS)(outs, dist)
    grads = torch.autograd.grad(val, outs[1:])
    out["Consistency"] = {"value": float(val), "grad_l2": float(torch.cat([t.reshape(-1) for t in grads]).norm())}
    out["TV"] = {"value": float(L.tv_loss(torch.randn(16, 20, 2, generator=g)))}
    return out


def main():
    from oracle import ref_shims
    R = ref_shims.load("metrics.losses")
    path = os.path.join(ROOT, "tests", "golden", "dropin_losses.json")
    with open(path, "w") as f:
        json.dump(evaluate(R), f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
