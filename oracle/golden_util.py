"""Shared case table + digest helpers for the golden fixtures  --  TEST INFRASTRUCTURE ONLY.

A *case* fully determines inputs from integer seeds (torch CPU generators are bit-stable for a
fixed torch build, and the GPU box runs this same image), so the committed fixtures only hold
digests of the reference's results: a handful of raw values plus norms per tensor.
"""
from __future__ import annotations

import json
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ENC_GAUSS = {"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3}
ENC_NONE = {"embedding": "none", "scale": 4, "embedding_size": 256, "coordinates_size": 3}
ENC_LOGF = {"embedding": "LogF", "scale": 4, "embedding_size": 256, "coordinates_size": 3}     # 42 frequencies x 3 coords x (sin, cos) = 252
NET_256 = {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256}
NET_WIRE = {"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256,
            "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}
NET_W2D = {"network_input_size": 3, "network_output_size": 2, "network_depth": 3, "network_width": 256,
           "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}
NET_MFN = {"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 512}
HDR_OPTS = {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}

# name -> (model, net, encoder, loss, loss_opts, batch, seed)
CASES = {
    "siren_l2":   ("SIREN", NET_256, ENC_GAUSS, "L2", None, 1000, 11),
    "siren_tanh": ("SIREN", dict(NET_256, last_tanh=True), ENC_GAUSS, "tanh", None, 777, 12),
    "siren_l1":   ("SIREN", NET_256, ENC_GAUSS, "L1", None, 640, 13),
    "ffn_l2":     ("FFN", NET_256, ENC_GAUSS, "L2", None, 1000, 14),
    "ffn_msle":   ("FFN", NET_256, ENC_GAUSS, "MSLE", None, 512, 15),
    "wire_hdr":   ("WIRE", NET_WIRE, ENC_NONE, "HDR", HDR_OPTS, 600, 16),
    "wire_l2":    ("WIRE", NET_WIRE, ENC_NONE, "L2", None, 500, 17),
    "fourier_l2": ("Fourier", NET_MFN, ENC_GAUSS, "L2", None, 300, 18),
    "gabor_tanh": ("Gabor", NET_MFN, ENC_GAUSS, "tanh", None, 300, 19),
    "wire2d_l2":  ("WIRE2D", NET_W2D, ENC_NONE, "L2", None, 400, 20),
    "wire2d_tanh": ("WIRE2D", dict(NET_W2D, last_tanh=True), ENC_NONE, "tanh", None, 384, 21),
    "siren_logf": ("SIREN", dict(NET_256, network_input_size=252), ENC_LOGF, "L2", None, 600, 22),
}
N_ADAM_STEPS = 3
LR = 5e-4


def case_inputs(name):
    """coords in [-1,1]^3 (coil, kx, ky as create_coords gives, src/data/utils.py:98-108) and a
    smooth complex target in [0,1] / small-magnitude k-space-like values."""
    model, net, enc, loss, opts, batch, seed = CASES[name]
    g = torch.Generator().manual_seed(1000 + seed)
    coords = torch.rand(batch, 3, generator=g) * 2 - 1
    if loss in ("MSLE",):
        gt = torch.rand(batch, 2, generator=g)
    elif loss in ("HDR",):
        gt = torch.randn(batch, 2, generator=g) * 0.05
    else:
        gt = torch.rand(batch, 2, generator=g) * 0.8 + 0.1
    return coords, gt


def case_mask(name):
    """Undersampling row mask (src/train.py:172-177).  WIRE+HDR only works in the reference when
    a mask is present: ``output.real`` is non-contiguous and ``view_as_complex``
    (src/metrics/losses.py:245) rejects it unless the boolean row-select made a copy first."""
    model, net, enc, loss, opts, batch, seed = CASES[name]
    if loss != "HDR":
        return None
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.rand(batch, generator=g) < 0.5


def tensor_digest(t: torch.Tensor, n_head=6):
    t = torch.view_as_real(t) if t.is_complex() else t
    f = t.detach().double().flatten()
    return {"n": int(f.numel()), "sum": float(f.sum()), "l2": float(f.norm()),
            "head": [float(v) for v in f[:n_head]], "tail": [float(v) for v in f[-n_head:]]}


def digest_close(a, b, rtol, atol_scale=1.0):
    """Compare two tensor digests; returns list of failure strings."""
    fails = []
    if a["n"] != b["n"]:
        return [f"numel {a['n']} != {b['n']}"]
    scale = max(a["l2"], 1e-30)
    # l2 norms agree relatively; raw values agree to rtol * rms
    if abs(a["l2"] - b["l2"]) > rtol * scale:
        fails.append(f"l2 {a['l2']} vs {b['l2']}")
    rms = scale / max(a["n"], 1) ** 0.5
    for key in ("head", "tail"):
        for x, y in zip(a[key], b[key]):
            if abs(x - y) > rtol * max(abs(x), rms * atol_scale):
                fails.append(f"{key} {x} vs {y}")
    return fails


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


# ---- loss-only unit cases: seeded [m,2] prediction / target, [bs_k,3] kcoords -----------------
LOSS_CASES = {
    "L2": None, "L1": None, "MSLE": None, "tanh": None,
    "HDR": HDR_OPTS, "HDR_nofilter": {"hdr_eps": 3e-3, "hdr_ff_sigma": 2.0, "hdr_ff_factor": 0.0},
    "LSL": {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.0},
    "TV": None, "Consistency": None,
}
LOSS_M, LOSS_BSK = 96, 192      # m rows enter the loss, bs_k rows of (unmasked) kcoords
TV_HW = (8, 12)
CONS_BOUNDS = [(0, 0.3), (0, 0.6), (0, 0.9), (0, 5)]


def loss_case_inputs(kind):
    g = torch.Generator().manual_seed(4242 + sum(map(ord, kind)))
    m = TV_HW[0] * TV_HW[1] if kind in ("TV", "Consistency") else LOSS_M
    out = torch.randn(m, 2, generator=g) * 0.3
    gt = torch.randn(m, 2, generator=g) * 0.3
    if kind == "MSLE":
        out, gt = out.abs(), gt.abs()
    kc = torch.rand(LOSS_BSK, 3, generator=g) * 2 - 1
    extra = [torch.randn(m, 2, generator=g) * 0.3 for _ in range(3)]
    dist = torch.rand(m, generator=g) * 1.4
    return out, gt, kc, extra, dist
