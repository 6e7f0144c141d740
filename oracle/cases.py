"""Run a golden case through the ORACLE restatement  --  TEST INFRASTRUCTURE ONLY.

``run_oracle_case`` mirrors ``make_golden.run_reference_case`` step for step but uses nothing from
/root/reference: oracle init (same RNG call order), oracle forward, closed-form loss gradients,
autograd only for the chain rule through the oracle's own forward, explicit Adam.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import golden_util as G
from . import inr_oracle as O


def case_setup(name):
    """(model_kind, net, enc_cfg, loss, opts, sd, encB, coords, gt, mask) with reference RNG order:
    torch.manual_seed(seed) -> encoder B -> model parameters."""
    model_kind, net, enc_cfg, loss_kind, opts, batch, seed = G.CASES[name]
    torch.manual_seed(seed)
    encB = O.encoder_init(enc_cfg)
    sd = O.MODEL_INIT[model_kind](dict(net))
    coords, gt = G.case_inputs(name)
    return model_kind, dict(net), enc_cfg, loss_kind, opts, sd, encB, coords, gt, G.case_mask(name)


def loss_and_grad(loss_kind, opts, out, gt, coords):
    """Closed-form value and d/d(out) as the reference training loop weights them
    (src/train.py:178-182)."""
    if loss_kind == "HDR":
        val, g, _ = O.loss_hdr(out, gt, coords, float(opts["hdr_ff_sigma"]), float(opts["hdr_eps"]),
                               float(opts["hdr_ff_factor"]))
        return val, g
    if loss_kind == "LSL":
        return O.loss_logspace(out, gt, float(opts["hdr_eps"]))
    return O.LOSS_TRAIN[loss_kind](out, gt)


def _cast(t, dtype):
    if t is None or dtype == torch.float32:
        return t
    return t.to(torch.complex128 if t.is_complex() else dtype)


def run_oracle_case(name, dtype=torch.float32):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt_full, mask = case_setup(name)
    sd = OrderedDict((k, _cast(v, dtype)) for k, v in sd.items())
    encB, coords, gt_full = _cast(encB, dtype), _cast(coords, dtype), _cast(gt_full, dtype)
    d = {"case": name, "init": {k: G.tensor_digest(v) for k, v in sd.items()},
         "encB": None if encB is None else G.tensor_digest(encB)}
    frozen = {k for k in sd if k.endswith("omega_0") or k.endswith("scale_0")}
    params = OrderedDict((k, v.clone().requires_grad_(k not in frozen)) for k, v in sd.items())
    live_keys = [k for k in params if k not in frozen]
    state = {}
    for k in live_keys:
        p = params[k]
        shape = torch.view_as_real(p).shape if p.is_complex() else p.shape
        state[k] = (torch.zeros(shape, dtype=dtype), torch.zeros(shape, dtype=dtype))
    d["losses"] = []
    for step in range(G.N_ADAM_STEPS):
        x = O.encode(coords, encB, enc_cfg["embedding"])
        out = O.model_forward(model_kind, params, x, net)
        gt = gt_full
        sel_out = out
        if mask is not None:
            sel_out, gt = out[mask], gt_full[mask]
        val, g = loss_and_grad(loss_kind, opts, sel_out.detach(), gt, coords)
        grads = torch.autograd.grad(sel_out, [params[k] for k in live_keys], grad_outputs=g)
        if step == 0:
            d["out"] = G.tensor_digest(sel_out)
            d["loss"] = float(val)
            d["grads"] = {k: G.tensor_digest(gr) for k, gr in zip(live_keys, grads)}
        d["losses"].append(float(val))
        with torch.no_grad():
            for k, gr in zip(live_keys, grads):
                p = params[k]
                pr = torch.view_as_real(p) if p.is_complex() else p
                grr = torch.view_as_real(gr.contiguous()) if gr.is_complex() else gr
                m, v = state[k]
                O.adam_step(pr, grr, m, v, step + 1, G.LR)
    d["final"] = {k: G.tensor_digest(v) for k, v in params.items()}
    return d
