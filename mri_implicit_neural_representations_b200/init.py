"""Parameter initialisation with the reference's distributions, drawn on the CPU with the same torch calls
in the same order so that torch.manual_seed(s) reproduces the reference's starting point bit for bit
(reference src/models/networks.py:13,57-62,79-89)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def encoder_matrix(enc: dict):
    if enc.get("embedding", "none") == "gauss":
        return torch.randn((enc["embedding_size"], enc["coordinates_size"])) * enc["scale"]
    return None


def chain_tensors(model: str, net: dict):
    """[(name, tensor)] in reference state_dict order for SIREN / FFN."""
    depth, width = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    dims = [(fin, width)] + [(width, width)] * (depth - 2) + [(width, fout)]
    out = []
    for i, (a, b) in enumerate(dims):
        lin = nn.Linear(a, b)
        if model == "SIREN":
            bound = 1.0 / a if i == 0 else math.sqrt(6.0 / a) / 30.0
            with torch.no_grad():
                lin.weight.uniform_(-bound, bound)
            prefix = f"model.{i}.linear"
        elif model == "FFN":
            prefix = f"model.{2 * i}"
        else:
            raise NotImplementedError(model)
        out.append((prefix + ".weight", lin.weight.detach().clone()))
        out.append((prefix + ".bias", lin.bias.detach().clone()))
    return out


def wire_tensors(net: dict):
    """[(name, tensor)] in reference state_dict order for WIRE (reference src/models/networks.py:206-252): frozen
    omega_0 / scale_0 scalars, a real first layer, `depth` complex hidden layers of width int(width/sqrt(2)), a complex
    final linear.  nn.Linear(dtype=cfloat) default init, same RNG order."""
    import numpy as np
    depth = net["network_depth"]
    hid = int(net["network_width"] / np.sqrt(2))
    fin, fout = net["network_input_size"], net["network_output_size"]
    out = []
    specs = [(fin, hid, torch.float, net["first_omega_0"])] + [(hid, hid, torch.cfloat, net["hidden_omega_0"])] * depth
    for i, (a, b, dt, om) in enumerate(specs):
        out.append((f"net.{i}.omega_0", om * torch.ones(1)))
        out.append((f"net.{i}.scale_0", net["scale"] * torch.ones(1)))
        lin = nn.Linear(a, b, dtype=dt)
        out.append((f"net.{i}.linear.weight", lin.weight.detach().clone()))
        out.append((f"net.{i}.linear.bias", lin.bias.detach().clone()))
    lin = nn.Linear(hid, fout, dtype=torch.cfloat)
    out.append((f"net.{depth + 1}.weight", lin.weight.detach().clone()))
    out.append((f"net.{depth + 1}.bias", lin.bias.detach().clone()))
    return out


def wire2d_tensors(net: dict):
    """[(name, tensor)] in reference state_dict order for WIRE2D (reference src/models/wire2d.py:62-104): per layer the
    frozen omega_0 / scale_0, `linear` and `scale_orth` (real on the first layer, complex after; width NOT reduced), a
    complex final linear.  Same nn.Linear calls in the same RNG order."""
    depth, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    out = []
    specs = [(fin, hid, torch.float, net["first_omega_0"])] + [(hid, hid, torch.cfloat, net["hidden_omega_0"])] * depth
    for i, (a, b, dt, om) in enumerate(specs):
        out.append((f"net.{i}.omega_0", om * torch.ones(1)))
        out.append((f"net.{i}.scale_0", net["scale"] * torch.ones(1)))
        for name in ("linear", "scale_orth"):
            lin = nn.Linear(a, b, dtype=dt)
            out.append((f"net.{i}.{name}.weight", lin.weight.detach().clone()))
            out.append((f"net.{i}.{name}.bias", lin.bias.detach().clone()))
    lin = nn.Linear(hid, fout, dtype=torch.cfloat)
    out.append((f"net.{depth + 1}.weight", lin.weight.detach().clone()))
    out.append((f"net.{depth + 1}.bias", lin.bias.detach().clone()))
    return out


def mfn_tensors(model: str, net: dict, input_scale: float = 2.0, weight_scale: float = 1.0, alpha: float = 6.0,
                beta: float = 1.0):
    """[(name, tensor)] in reference state_dict order for FourierNet ('Fourier'), GaborNet / KGaborNet ('Gabor',
    'KGabor') and the multiscale variants ('MultiscaleFourier', 'BoundedFourier'), same torch calls in the same order
    as reference src/models/mfn.py (:15-30 MFNBase, :50-55 FourierLayer, :61-83, :102-113 GaborLayer, :134-162,
    :216-251, :300-340)."""
    L, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    lins = [nn.Linear(hid, hid, True) for _ in range(L)]
    out_lin = nn.Linear(hid, fout)
    b = math.sqrt(weight_scale / hid)
    for lin in lins:
        lin.weight.data.uniform_(-b, b)
    if model in ("Gabor", "KGabor"):
        res = []
        for i, l in enumerate(lins):
            res += [(f"linear.{i}.weight", l.weight.detach().clone()), (f"linear.{i}.bias", l.bias.detach().clone())]
        res += [("output_linear.weight", out_lin.weight.detach().clone()), ("output_linear.bias", out_lin.bias.detach().clone())]
        ws = input_scale / math.sqrt(L + 1)
        for i in range(L + 1):
            lin = nn.Linear(fin, hid)                                               # GaborLayer.__init__, :104
            mu = 2 * torch.rand(hid, fin) - 1                                       # :109
            gamma = torch.distributions.gamma.Gamma(alpha / (L + 1), beta).sample((hid,))   # :110-112
            lin.weight.data *= ws * torch.sqrt(gamma[:, None])                      # :113
            lin.bias.data.uniform_(-math.pi, math.pi)                               # :114
            res += [(f"filters.{i}.mu", mu), (f"filters.{i}.gamma", gamma),
                    (f"filters.{i}.linear.weight", lin.weight.detach().clone()),
                    (f"filters.{i}.linear.bias", lin.bias.detach().clone())]
        return res
    multi = model != "Fourier"
    if model == "BoundedFourier":                       # BoundedLinear replaces the scaled-uniform linears (default init)
        lins = [nn.Linear(hid, hid, True) for _ in range(L)]
    fscale = (weight_scale if multi else input_scale) / math.sqrt(L + 1)
    filt = []
    for _ in range(L + 1):
        f = nn.Linear(fin, hid)
        f.weight.data *= fscale
        f.bias.data.uniform_(-math.pi, math.pi)
        filt.append(f)
    outs = [nn.Linear(hid, fout) for _ in range(L + 1)] if multi else None
    res = []
    for i, l in enumerate(lins):
        pre = f"linear.{i}.linear" if model == "BoundedFourier" else f"linear.{i}"
        res += [(pre + ".weight", l.weight.detach().clone()), (pre + ".bias", l.bias.detach().clone())]
    if multi:
        for i, o in enumerate(outs):
            res += [(f"output_linear.{i}.weight", o.weight.detach().clone()), (f"output_linear.{i}.bias", o.bias.detach().clone())]
    else:
        res += [("output_linear.weight", out_lin.weight.detach().clone()), ("output_linear.bias", out_lin.bias.detach().clone())]
    for i, f in enumerate(filt):
        res += [(f"filters.{i}.linear.weight", f.weight.detach().clone()), (f"filters.{i}.linear.bias", f.bias.detach().clone())]
    return res
