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
