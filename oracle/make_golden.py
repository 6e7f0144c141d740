"""Generate tests/golden/*.json from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For every case in ``golden_util.CASES`` the reference's own modules (model class, loss class,
``torch.optim.Adam``) run the loop body of src/train.py:158-192 for N_ADAM_STEPS steps on the
seeded inputs; digests of the initial parameters, first-step output / loss / gradients and the
parameters after the last step are stored.  The fixtures are what the oracle restatement (and,
through it, the CUDA path) is pinned against on machines where /root/reference does not exist.
"""
from __future__ import annotations

import json
import os

import torch

from . import golden_util as G
from . import ref_shims


def run_reference_case(name, dtype=torch.float32):
    N, M, W2, L = ref_shims.load("models.networks", "models.mfn", "models.wire2d", "metrics.losses")
    model_kind, net, enc_cfg, loss_kind, opts, batch, seed = G.CASES[name]
    torch.manual_seed(seed)
    enc = N.Positional_Encoder(enc_cfg, device="cpu")
    cls = {"SIREN": N.SIREN, "FFN": N.FFN, "WIRE": N.WIRE, "WIRE2D": W2.WIRE2D,
           "Fourier": M.FourierNet, "Gabor": M.GaborNet}[model_kind]
    model = cls(dict(net))
    model.train()
    if dtype == torch.float64:                    # nn.Module.double() leaves complex64 params alone
        for p in model.parameters():
            p.data = p.data.to(torch.complex128 if p.is_complex() else torch.float64)
        if enc.B is not None:
            enc.B = enc.B.double()
    optim = torch.optim.Adam(model.parameters(), lr=G.LR, betas=(0.9, 0.999), weight_decay=0.0)
    if loss_kind == "L2":
        loss_fn = torch.nn.MSELoss()
    elif loss_kind == "L1":
        loss_fn = torch.nn.L1Loss()
    elif loss_kind == "MSLE":
        loss_fn = L.MSLELoss()
    elif loss_kind == "tanh":
        loss_fn = L.TanhL2Loss()
    elif loss_kind == "HDR":
        loss_fn = L.HDRLoss_FF(opts)
    coords, gt_full = G.case_inputs(name)
    coords, gt_full = coords.to(dtype), gt_full.to(dtype)
    mask = G.case_mask(name)
    out_d = {"case": name, "torch": torch.__version__,
             "init": {k: G.tensor_digest(v) for k, v in model.state_dict().items()},
             "encB": None if enc.B is None else G.tensor_digest(enc.B)}
    for step in range(G.N_ADAM_STEPS):
        x = enc.embedding(coords)
        out = model(x)
        optim.zero_grad()
        gt = gt_full
        if mask is not None:                      # src/train.py:172-177
            out = out[mask]
            gt = gt_full[mask]
        if loss_kind == "HDR" and dtype == torch.float64:
            # src/metrics/losses.py:244 only builds the complex view for float32 inputs
            loss, _ = loss_fn(torch.view_as_complex(out.contiguous()),
                              torch.view_as_complex(gt.contiguous()), coords)
        elif loss_kind in ("HDR", "tanh"):        # src/train.py:178-182
            loss, _ = loss_fn(out, gt, coords)
        else:
            loss = 0.5 * loss_fn(out, gt)
        loss.backward()
        if step == 0:
            out_d["out"] = G.tensor_digest(out)
            out_d["loss"] = float(loss)
            out_d["grads"] = {k: G.tensor_digest(p.grad) for k, p in model.named_parameters()
                              if p.grad is not None}
        optim.step()
        out_d.setdefault("losses", []).append(float(loss))
    out_d["final"] = {k: G.tensor_digest(v) for k, v in model.state_dict().items()}
    return out_d


def run_reference_losses():
    """Reference loss classes + autograd on seeded tensors, weighted as the training loops do
    (src/train.py:178-182, src/train_kspace_multiscale.py:173-190)."""
    L = ref_shims.load("metrics.losses")
    res = {}
    for kind, opts in G.LOSS_CASES.items():
        out, gt, kc, extra, dist = G.loss_case_inputs(kind)
        out = out.clone().requires_grad_(True)
        if kind == "L2":
            val = 0.5 * torch.nn.MSELoss()(out, gt)
        elif kind == "L1":
            val = 0.5 * torch.nn.L1Loss()(out, gt)
        elif kind == "MSLE":
            val = 0.5 * L.MSLELoss()(out, gt)
        elif kind == "tanh":
            val, _ = L.TanhL2Loss()(out, gt, kc)
        elif kind.startswith("HDR"):
            val, _ = L.HDRLoss_FF(opts)(out, gt, kc)
        elif kind == "LSL":
            val = 0.5 * L.LogSpaceLoss(opts)(out, gt)
        elif kind == "TV":
            val = L.tv_loss(out.view(G.TV_HW[0], G.TV_HW[1], 2))
        elif kind == "Consistency":
            outs = [out] + [e.clone().requires_grad_(True) for e in extra]
            val = 0.1 * L.ConsistencyLoss(G.CONS_BOUNDS)(outs, dist)
            val.backward()
            res[kind] = {"value": float(val), "grads": [G.tensor_digest(
                o.grad if o.grad is not None else torch.zeros_like(o), n_head=12) for o in outs]}
            continue
        val.backward()
        res[kind] = {"value": float(val), "dout": G.tensor_digest(out.grad, n_head=12)}
    return res


def main():
    """python -m oracle.make_golden [case ...]: regenerate everything, or only the named model cases."""
    import sys
    only = sys.argv[1:]
    os.makedirs(G.GOLDEN_DIR, exist_ok=True)
    if not only:
        with open(os.path.join(G.GOLDEN_DIR, "losses.json"), "w") as f:
            json.dump(run_reference_losses(), f, indent=1)
    for name in (only or G.CASES):
        d = run_reference_case(name)
        d["fp64"] = run_reference_case(name, torch.float64)
        with open(os.path.join(G.GOLDEN_DIR, name + ".json"), "w") as f:
            json.dump(d, f, indent=1)
        print("wrote", name, "loss", d["loss"])


if __name__ == "__main__":
    main()
