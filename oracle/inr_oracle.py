"""CPU oracle for the INR fitting hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (plain torch-on-CPU tensor algebra, explicit
backward formulas where the CUDA kernels implement explicit formulas) of the arithmetic
the reference performs per training batch.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the product
package never does (it fails loudly without its CUDA library instead).

Parity pin: every function here is checked against the *unmodified* reference modules
imported from /root/reference (``oracle/ref_shims.py``) by ``tests/test_oracle_vs_reference.py``
(runs wherever /root/reference exists) and against the committed digests under
``tests/golden/`` that ``oracle/make_golden.py`` produced from those reference modules.

Third-party pieces the reference calls (fastmri 0.3.0 fft2c / ifft2c / complex_abs / rss, scikit-image 0.18.1
structural_similarity) are restated here from the libraries' published definitions; the libraries are not in this
image, so parity with them is UNPINNED (checked against independent constructions in
tests/test_third_party_restatements.py).

Reference citations are relative to /root/reference/.
All functions are dtype-generic (float32 mirrors the reference, float64 is used as a
tighter yardstick in some tests).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

TWO_PI = 2.0 * math.pi

# --------------------------------------------------------------------------------------
# Positional encoder  (src/models/networks.py:7-35)
# --------------------------------------------------------------------------------------

def encoder_init(enc_cfg: dict):
    """B matrix exactly as src/models/networks.py:12-17 draws it (CPU RNG, same call)."""
    kind = enc_cfg["embedding"]
    if kind == "gauss":
        return torch.randn((enc_cfg["embedding_size"], enc_cfg["coordinates_size"])) * enc_cfg["scale"]
    if kind == "LogF":
        steps = int(enc_cfg["embedding_size"] / (2 * enc_cfg["coordinates_size"]))
        return (2.0 ** torch.linspace(0.0, enc_cfg["scale"], steps=steps)).reshape(-1, 1)
    if kind == "none":
        return None
    raise NotImplementedError(kind)


def encode(x: torch.Tensor, B, kind: str) -> torch.Tensor:
    """gamma(x) of src/models/networks.py:23-35."""
    if kind == "LogF":
        parts = []
        for c in range(3):
            arg = (TWO_PI * x[:, c:c + 1]) @ B.T
            parts += [torch.sin(arg), torch.cos(arg)]
        return torch.cat(parts, dim=-1)
    if B is not None:
        arg = (TWO_PI * x) @ B.t()
        return torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    return x


# --------------------------------------------------------------------------------------
# Parameter initialisation: same torch calls in the same order as the reference so that
# torch.manual_seed(s) gives bit-identical starts.
# --------------------------------------------------------------------------------------

def _lin(sd, prefix, lin: nn.Linear):
    sd[prefix + ".weight"] = lin.weight.detach().clone()
    if lin.bias is not None:
        sd[prefix + ".bias"] = lin.bias.detach().clone()


def siren_init(net: dict) -> "OrderedDict[str, torch.Tensor]":
    """src/models/networks.py:74-119."""
    depth, width = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    dims = [(fin, width)] + [(width, width)] * (depth - 2) + [(width, fout)]
    sd = OrderedDict()
    for i, (a, b) in enumerate(dims):
        lin = nn.Linear(a, b)
        bound = 1.0 / a if i == 0 else math.sqrt(6.0 / a) / 30.0   # :85-89 (w0 == 30 always)
        with torch.no_grad():
            lin.weight.uniform_(-bound, bound)
        _lin(sd, f"model.{i}.linear", lin)
    return sd


def ffn_init(net: dict):
    """src/models/networks.py:48-65 (nn.Sequential indices 0,2,4,... hold the Linears)."""
    depth, width = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    dims = [(fin, width)] + [(width, width)] * (depth - 2) + [(width, fout)]
    sd = OrderedDict()
    for i, (a, b) in enumerate(dims):
        _lin(sd, f"model.{2 * i}", nn.Linear(a, b))
    return sd


def wire_init(net: dict):
    """src/models/networks.py:206-252."""
    depth = net["network_depth"]
    hid = int(net["network_width"] / np.sqrt(2))
    fin, fout = net["network_input_size"], net["network_output_size"]
    sd = OrderedDict()
    specs = [(fin, hid, torch.float, net["first_omega_0"])] + \
            [(hid, hid, torch.cfloat, net["hidden_omega_0"])] * depth
    for i, (a, b, dt, om) in enumerate(specs):
        sd[f"net.{i}.omega_0"] = om * torch.ones(1)
        sd[f"net.{i}.scale_0"] = net["scale"] * torch.ones(1)
        _lin(sd, f"net.{i}.linear", nn.Linear(a, b, dtype=dt))
    _lin(sd, f"net.{depth + 1}", nn.Linear(hid, fout, dtype=torch.cfloat))
    return sd


def wire2d_init(net: dict):
    """src/models/wire2d.py:62-104."""
    depth, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    sd = OrderedDict()
    specs = [(fin, hid, torch.float, net["first_omega_0"])] + \
            [(hid, hid, torch.cfloat, net["hidden_omega_0"])] * depth
    for i, (a, b, dt, om) in enumerate(specs):
        sd[f"net.{i}.omega_0"] = om * torch.ones(1)
        sd[f"net.{i}.scale_0"] = net["scale"] * torch.ones(1)
        _lin(sd, f"net.{i}.linear", nn.Linear(a, b, dtype=dt))
        _lin(sd, f"net.{i}.scale_orth", nn.Linear(a, b, dtype=dt))
    _lin(sd, f"net.{depth + 1}", nn.Linear(hid, fout, dtype=torch.cfloat))
    return sd


def _mfn_base_init(sd, hidden, out, n_layers, weight_scale, bias=True):
    """src/models/mfn.py:15-30 (RNG order: linears, output_linear, then the uniform_ redraws)."""
    lins = [nn.Linear(hidden, hidden, bias) for _ in range(n_layers)]
    out_lin = nn.Linear(hidden, out)
    b = math.sqrt(weight_scale / hidden)
    for lin in lins:
        lin.weight.data.uniform_(-b, b)
    return lins, out_lin


def _fourier_filter(fin, hidden, scale):
    """src/models/mfn.py:50-55."""
    lin = nn.Linear(fin, hidden)
    lin.weight.data *= scale
    lin.bias.data.uniform_(-math.pi, math.pi)
    return lin


def fourier_init(net: dict, input_scale=2.0, weight_scale=1.0):
    """FourierNet, src/models/mfn.py:61-83; state_dict order follows module registration."""
    L, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    sd = OrderedDict()
    lins, out_lin = _mfn_base_init(sd, hid, fout, L, weight_scale)
    filt = [_fourier_filter(fin, hid, input_scale / math.sqrt(L + 1)) for _ in range(L + 1)]
    for i, l in enumerate(lins):
        _lin(sd, f"linear.{i}", l)
    _lin(sd, "output_linear", out_lin)
    for i, f in enumerate(filt):
        _lin(sd, f"filters.{i}.linear", f)
    return sd


def gabor_init(net: dict, input_scale=2.0, weight_scale=1.0, alpha=6.0, beta=1.0):
    """GaborNet / KGaborNet, src/models/mfn.py:96-113,133-162."""
    L, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    sd = OrderedDict()
    lins, out_lin = _mfn_base_init(sd, hid, fout, L, weight_scale)
    for i, l in enumerate(lins):
        _lin(sd, f"linear.{i}", l)
    _lin(sd, "output_linear", out_lin)
    ws = input_scale / math.sqrt(L + 1)
    for i in range(L + 1):
        lin = nn.Linear(fin, hid)
        mu = 2 * torch.rand(hid, fin) - 1
        gamma = torch.distributions.gamma.Gamma(alpha / (L + 1), beta).sample((hid,))
        lin.weight.data *= ws * torch.sqrt(gamma[:, None])
        lin.bias.data.uniform_(-math.pi, math.pi)
        sd[f"filters.{i}.mu"] = mu
        sd[f"filters.{i}.gamma"] = gamma
        _lin(sd, f"filters.{i}.linear", lin)
    return sd


def multiscale_init(net: dict, bounded: bool, weight_scale=1.0):
    """MultiscaleKFourier / MultiscaleBoundedFourier, src/models/mfn.py:206-253,288-342.

    The bounded variant first builds (and discards) the MFNBase linears, consuming RNG,
    then replaces them by default-initialised BoundedLinear layers (:316-326)."""
    L, hid = net["network_depth"], net["network_width"]
    fin, fout = net["network_input_size"], net["network_output_size"]
    sd = OrderedDict()
    lins, _single_out = _mfn_base_init(sd, hid, fout, L, weight_scale)
    if bounded:
        lins = [nn.Linear(hid, hid, True) for _ in range(L)]
    filt = [_fourier_filter(fin, hid, weight_scale / math.sqrt(L + 1)) for _ in range(L + 1)]
    outs = [nn.Linear(hid, fout) for _ in range(L + 1)]
    for i, l in enumerate(lins):
        _lin(sd, f"linear.{i}.linear" if bounded else f"linear.{i}", l)
    # registration order in the reference: linear, output_linear (re-assigned in place), filters
    for i, o in enumerate(outs):
        _lin(sd, f"output_linear.{i}", o)
    for i, f in enumerate(filt):
        _lin(sd, f"filters.{i}.linear", f)
    return sd


# --------------------------------------------------------------------------------------
# Forward passes (functional, on a state_dict)
# --------------------------------------------------------------------------------------

def _affine(x, w, b=None):
    y = x @ w.t().to(x.dtype) if not w.is_complex() or x.is_complex() else x.to(w.dtype) @ w.t()
    return y if b is None else y + b


def siren_forward(sd, x, depth, last_tanh=False, last_linear=True, w0=30.0, trace=None):
    """src/models/networks.py:91-96,121-124.  ``trace`` (list) receives (z, h) per layer.
    The reference builds first + (depth - 2) hidden + last layers (:111-115), i.e. TWO layers for depth 1 as for depth 2."""
    h = x
    depth = max(int(depth), 2)
    for i in range(depth):
        z = h @ sd[f"model.{i}.linear.weight"].t() + sd[f"model.{i}.linear.bias"]
        if i == depth - 1:
            if last_tanh:
                h = torch.tanh(z)
            elif last_linear:
                h = z
            else:
                h = torch.sin(w0 * z)
        else:
            h = torch.sin(w0 * z)
        if trace is not None:
            trace.append((z, h))
    return h


def ffn_forward(sd, x, depth, trace=None):
    """src/models/networks.py:57-69: ReLU hidden layers, Sigmoid on the output."""
    h = x
    for i in range(depth):
        z = h @ sd[f"model.{2 * i}.weight"].t() + sd[f"model.{2 * i}.bias"]
        h = torch.sigmoid(z) if i == depth - 1 else torch.relu(z)
        if trace is not None:
            trace.append((z, h))
    return h


def gabor_act(z, omega, sigma):
    """exp(1j*omega*z - |sigma*z|^2), src/models/networks.py:199-204, written out in re/im."""
    if z.is_complex():
        a, b = z.real, z.imag
    else:
        a, b = z, torch.zeros_like(z)
    mag = torch.exp(-omega * b - (sigma * sigma) * (a * a + b * b))
    return torch.complex(mag * torch.cos(omega * a), mag * torch.sin(omega * a))


def wire_forward(sd, x, depth, trace=None):
    """src/models/networks.py:254-258; depth hidden complex layers after the real first one."""
    h = x
    for i in range(depth + 1):
        w, b = sd[f"net.{i}.linear.weight"], sd[f"net.{i}.linear.bias"]
        z = (h @ w.t() + b) if i == 0 else (h @ w.t() + b)
        h = gabor_act(z, sd[f"net.{i}.omega_0"].to(x.dtype), sd[f"net.{i}.scale_0"].to(x.dtype))
        if trace is not None:
            trace.append((z, h))
    w, b = sd[f"net.{depth + 1}.weight"], sd[f"net.{depth + 1}.bias"]
    out = h @ w.t() + b
    if trace is not None:
        trace.append((out, out.real))
    return out.real


def wire2d_forward(sd, x, depth, trace=None, last_tanh=False):
    """src/models/wire2d.py:49-60,112-118; last_tanh (:106-107): torch.nn.Tanh on the complex output, then .real."""
    h = x
    for i in range(depth + 1):
        l = h @ sd[f"net.{i}.linear.weight"].t() + sd[f"net.{i}.linear.bias"]
        s = h @ sd[f"net.{i}.scale_orth.weight"].t() + sd[f"net.{i}.scale_orth.bias"]
        om = sd[f"net.{i}.omega_0"].to(x.dtype)
        sg = sd[f"net.{i}.scale_0"].to(x.dtype)
        la, lb = (l.real, l.imag) if l.is_complex() else (l, torch.zeros_like(l))
        sa, sb = (s.real, s.imag) if s.is_complex() else (s, torch.zeros_like(s))
        mag = torch.exp(-om * lb - sg * sg * (la * la + lb * lb + sa * sa + sb * sb))
        h = torch.complex(mag * torch.cos(om * la), mag * torch.sin(om * la))
        if trace is not None:
            trace.append((l, s, h))
    out = h @ sd[f"net.{depth + 1}.weight"].t() + sd[f"net.{depth + 1}.bias"]
    if last_tanh:
        out = torch.tanh(out)
    return out.real


def _filter(sd, i, x, gabor):
    p = x @ sd[f"filters.{i}.linear.weight"].t() + sd[f"filters.{i}.linear.bias"]
    g = torch.sin(p)
    if gabor:   # src/models/mfn.py:126-131
        mu, gamma = sd[f"filters.{i}.mu"], sd[f"filters.{i}.gamma"]
        D = (x ** 2).sum(-1)[..., None] + (mu ** 2).sum(-1)[None, :] - 2 * x @ mu.T
        g = g * torch.exp(-0.5 * D * gamma[None, :])
    return g


def mfn_forward(sd, x, depth, gabor=False, trace=None):
    """FourierNet / GaborNet / KGaborNet forward, src/models/mfn.py:34-43,85-94,195-204."""
    z = _filter(sd, 0, x, gabor)
    if trace is not None:
        trace.append(z)
    for i in range(1, depth + 1):
        z = _filter(sd, i, x, gabor) * (z @ sd[f"linear.{i - 1}.weight"].t() + sd[f"linear.{i - 1}.bias"])
        if trace is not None:
            trace.append(z)
    return z @ sd["output_linear.weight"].t() + sd["output_linear.bias"]


def multiscale_forward(sd, x, depth, dist=None, boundaries=None, output_layers=(1, 3, 5, 7)):
    """src/models/mfn.py:255-267 (boundaries None) and :281-286,344-356 (bounded).

    ``dist`` 1-D [bs] zeroes whole rows; 2-D [bs,1] reproduces the reference's
    tuple-index quirk (only column 0 is zeroed)."""
    outs = []
    z = _filter(sd, 0, x, False)
    for i in range(1, depth + 1):
        zin = z
        if boundaries is not None:
            lo, hi = boundaries[i - 1]
            zin = z.clone()
            ind = torch.where((dist < lo) | (dist > hi))
            zin[ind] = 0
            w, b = sd[f"linear.{i - 1}.linear.weight"], sd[f"linear.{i - 1}.linear.bias"]
        else:
            w, b = sd[f"linear.{i - 1}.weight"], sd[f"linear.{i - 1}.bias"]
        z = _filter(sd, i, x, False) * (zin @ w.t() + b)
        if i in output_layers:
            outs.append(z @ sd[f"output_linear.{i}.weight"].t() + sd[f"output_linear.{i}.bias"])
    return outs


# --------------------------------------------------------------------------------------
# Losses: value and d(loss)/d(out) in closed form.
# ``train_weight`` reproduces src/train.py:178-182 (0.5 for L2/L1/MSLE, 1 otherwise).
# --------------------------------------------------------------------------------------

def loss_l2(out, gt):
    m = out.shape[0]
    e = out - gt
    return (e * e).sum() / (4 * m), e / (2 * m)


def loss_l1(out, gt):
    m = out.shape[0]
    e = out - gt
    return e.abs().sum() / (4 * m), torch.sign(e) / (4 * m)


def loss_msle(out, gt, eps=1e-9):
    """src/metrics/losses.py:18-27 times the 0.5 of src/train.py:182."""
    m = out.shape[0]
    lx, ly = torch.log(out + 1 + eps), torch.log(gt + 1 + eps)
    return ((lx - ly) ** 2).sum() / (4 * m), (lx - ly) / ((out + 1 + eps) * 2 * m)


def loss_tanh(out, gt):
    """src/metrics/losses.py:130-131 (with_mag False)."""
    m = out.shape[0]
    tx, ty = torch.tanh(out), torch.tanh(gt)
    return ((tx - ty) ** 2).sum() / (2 * m), (tx - ty) * (1 - tx * tx) / m


def loss_logspace(out, gt, eps, weight=0.5):
    """src/metrics/losses.py:214-223; weight 0.5 from src/train_kspace_multiscale.py:190."""
    m = out.shape[0]
    e = out - gt
    d = torch.sqrt((out * out).sum(-1, keepdim=True)) + eps
    val = weight * ((e * e).sum(-1) / d.squeeze(-1) ** 2).sum() / m
    return val, weight * 2 * e / (d * d * m)


def loss_hdr(out, gt, kcoords, sigma, eps, factor):
    """HDRLoss_FF in its exactly separable form (src/metrics/losses.py:236-262).

    reg broadcasts [bs_k,1] against [m] -> [bs_k,m]; its mean factorises into
    factor * mean_i((1-f_i)^2) * mean_j(|x_j|^2/(|x_j|+eps)^2)."""
    m = out.shape[0]
    e = out - gt
    ae = torch.sqrt((e * e).sum(-1))
    ax = torch.sqrt((out * out).sum(-1))
    d = ax + eps
    f = torch.exp(-(kcoords[:, 1] ** 2 + kcoords[:, 2] ** 2) / (2 * sigma ** 2))
    A = ((1 - f) ** 2).mean()
    lg = torch.log(ae / d)
    reg = factor * A * (ax * ax / (d * d)).sum() / m
    val = (lg * lg).sum() / m + reg
    grad = (2 * lg / (ae * ae * m))[:, None] * e + (factor * A * 2 / (d * d * m))[:, None] * out
    return val, grad, reg


def loss_hdr_reference_shape(out, gt, kcoords, sigma, eps, factor):
    """Literal [bs_k, m] formulation (small sizes only) used to validate ``loss_hdr``."""
    x = torch.view_as_complex(out.contiguous())
    y = torch.view_as_complex(gt.contiguous())
    f = torch.exp(-(kcoords[:, 1] ** 2 + kcoords[:, 2] ** 2) / (2 * sigma ** 2)).unsqueeze(-1)
    d = x.detach().abs() + eps
    loss = torch.log((x - y).abs() / d) ** 2
    reg = factor * ((x - x * f).abs() / d) ** 2
    return loss.mean() + reg.mean(), reg.mean()


def loss_consistency(outs, dist, bounds, weight=0.1):
    """src/metrics/losses.py:315-324 with 1-D ``dist``; gradient only into outs[i+1]."""
    total = outs[0].new_zeros(())
    grads = [torch.zeros_like(o) for o in outs]
    for i in range(len(bounds) - 1):
        lo, hi = bounds[i]
        sel = (dist < lo) | (dist > hi)
        n = int(sel.sum())
        if n:
            diff = (outs[i + 1] - outs[i])[sel]
            total = total + weight * (diff * diff).mean()
            grads[i + 1][sel] += weight * 2 * diff / (n * outs[i].shape[1])
    return total, grads


def loss_tv(out, H, W, weight=1e-4):
    """src/metrics/losses.py:326-343 on out.view(H, W, 2)."""
    img = out.view(H, W, 2)
    dh = img[:-1] - img[1:]
    dw = img[:, :-1] - img[:, 1:]
    val = weight * (dh.abs().mean() + dw.abs().mean())
    g = torch.zeros_like(img)
    sh = weight * torch.sign(dh) / dh.numel()
    sw = weight * torch.sign(dw) / dw.numel()
    g[:-1] += sh
    g[1:] -= sh
    g[:, :-1] += sw
    g[:, 1:] -= sw
    return val, g.reshape(-1, 2)


def reg_l1(params, lam):
    """src/models/regularization.py:21-28."""
    return lam * sum(p.abs().sum() for p in params)


def reg_l2(params, lam):
    """src/models/regularization.py:30-36."""
    return lam * abs(sum((p ** 2).sum() for p in params))


LOSS_TRAIN = {"L2": loss_l2, "L1": loss_l1, "MSLE": loss_msle, "tanh": loss_tanh}


# --------------------------------------------------------------------------------------
# Explicit backward for the real-valued chains (what the CUDA dgrad/wgrad kernels compute)
# --------------------------------------------------------------------------------------

def siren_backward(sd, x, trace, dout, depth, last_tanh=False, w0=30.0):
    """Manual backward of ``siren_forward``.  Returns
    (grads dict keyed like sd, list of dZ per layer)."""
    grads, dzs = {}, [None] * depth
    dh = dout
    for i in reversed(range(depth)):
        z, h = trace[i]
        if i == depth - 1:
            dz = dh * (1 - h * h) if last_tanh else dh
        else:
            dz = dh * (w0 * torch.cos(w0 * z))
        dzs[i] = dz
        hin = x if i == 0 else trace[i - 1][1]
        grads[f"model.{i}.linear.weight"] = dz.t() @ hin
        grads[f"model.{i}.linear.bias"] = dz.sum(0)
        dh = dz @ sd[f"model.{i}.linear.weight"]
    return grads, dzs


def ffn_backward(sd, x, trace, dout, depth, masks=None):
    """``masks`` (optional list of 0/1 tensors per hidden layer) teacher-forces the ReLU derivative: the sign of
    a pre-activation within rounding error of 0 is precision-dependent, so per-layer gradient parity is
    judged with the masks of the implementation under test."""
    grads, dzs = {}, [None] * depth
    dh = dout
    for i in reversed(range(depth)):
        z, h = trace[i]
        if i == depth - 1:
            dz = dh * h * (1 - h)
        else:
            dz = dh * ((z > 0).to(dh.dtype) if masks is None else masks[i].to(dh.dtype))
        dzs[i] = dz
        hin = x if i == 0 else trace[i - 1][1]
        grads[f"model.{2 * i}.weight"] = dz.t() @ hin
        grads[f"model.{2 * i}.bias"] = dz.sum(0)
        dh = dz @ sd[f"model.{2 * i}.weight"]
    return grads, dzs


# --------------------------------------------------------------------------------------
# Adam (torch.optim.Adam semantics, src/train.py:76) and the per-epoch lr schedule (:153,:251)
# --------------------------------------------------------------------------------------

def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One step for one flat real tensor; ``t`` is the 1-based step count.  In place."""
    if weight_decay:
        g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)
    return p


def lr_at_epoch(lr0, epoch, max_epoch):
    return lr0 * 0.2 ** min(epoch / max_epoch, 1)


# --------------------------------------------------------------------------------------
# Quality metrics (src/models/utils.py:227-250) and the fastmri pieces they go through
# --------------------------------------------------------------------------------------

def psnr(x, xhat, epsilon=1e-10):
    """10*log10(max(x)/(mse+eps)) -- note max is NOT squared (src/models/utils.py:236-250)."""
    return 10 * torch.log10(torch.max(x) / (torch.mean((x - xhat) ** 2) + epsilon))


def ifft2c(data):
    """fastmri==0.3.0 ifft2c: centred orthonormal 2-D inverse FFT on [..., H, W, 2]."""
    c = torch.view_as_complex(data.contiguous())
    c = torch.fft.ifftshift(c, dim=(-2, -1))
    c = torch.fft.ifftn(c, dim=(-2, -1), norm="ortho")
    c = torch.fft.fftshift(c, dim=(-2, -1))
    return torch.view_as_real(c)


def fft2c(data):
    c = torch.view_as_complex(data.contiguous())
    c = torch.fft.ifftshift(c, dim=(-2, -1))
    c = torch.fft.fftn(c, dim=(-2, -1), norm="ortho")
    c = torch.fft.fftshift(c, dim=(-2, -1))
    return torch.view_as_real(c)


def complex_abs(data):
    return (data ** 2).sum(dim=-1).sqrt()


def rss(data, dim=0):
    return torch.sqrt((data ** 2).sum(dim))


def ssim(x, xhat, win=7, K1=0.01, K2=0.03):
    """skimage==0.18.1 structural_similarity defaults (uniform 7x7 window, sample covariance,
    border crop) with the reference's data_range (src/models/utils.py:227-233)."""
    from scipy.ndimage import uniform_filter
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(xhat, dtype=np.float64)
    data_range = max(x.max(), y.max()) - min(x.min(), y.min())
    NP = win * win
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(x, win), uniform_filter(y, win)
    uxx, uyy, uxy = uniform_filter(x * x, win), uniform_filter(y * y, win), uniform_filter(x * y, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return S[pad:-pad, pad:-pad].mean()


# --------------------------------------------------------------------------------------
# Whole-step driver used by parity tests and by bench.py's cpu_baseline leg
# --------------------------------------------------------------------------------------

MODEL_INIT = {"SIREN": siren_init, "FFN": ffn_init, "WIRE": wire_init, "WIRE2D": wire2d_init,
              "Fourier": fourier_init, "Gabor": gabor_init, "KGabor": gabor_init}


def model_forward(kind, sd, x, net, trace=None):
    d = net["network_depth"]
    if kind == "SIREN":
        return siren_forward(sd, x, d, net.get("last_tanh", False), net.get("network_last_linear", True), trace=trace)
    if kind == "FFN":
        return ffn_forward(sd, x, d, trace=trace)
    if kind == "WIRE":
        return wire_forward(sd, x, d, trace=trace)
    if kind == "WIRE2D":
        return wire2d_forward(sd, x, d, trace=trace, last_tanh=net.get("last_tanh", False))
    if kind == "Fourier":
        return mfn_forward(sd, x, d, False, trace=trace)
    if kind in ("Gabor", "KGabor"):
        return mfn_forward(sd, x, d, True, trace=trace)
    raise NotImplementedError(kind)


def train_steps(kind, net, sd, encB, enc_kind, coords, gt, n_steps, batch, lr,
                loss="L2", loss_opts=None, betas=(0.9, 0.999), mask=None, tv=None):
    """Grid-order mini-batches (shuffle=False, src/models/utils.py:84-90) through
    forward -> loss -> autograd backward -> Adam, exactly the loop body of src/train.py:158-192
    restricted to fused-kernel territory.  Returns (losses, final sd)."""
    params = OrderedDict((k, v.clone().requires_grad_(v.is_floating_point() or v.is_complex()))
                         for k, v in sd.items())
    frozen = {k for k in params if k.endswith("omega_0") or k.endswith("scale_0")}
    for k in frozen:
        params[k].requires_grad_(False)
    state = {k: (torch.zeros_like(torch.view_as_real(p) if p.is_complex() else p),
                 torch.zeros_like(torch.view_as_real(p) if p.is_complex() else p))
             for k, p in params.items() if k not in frozen}
    N = coords.shape[0]
    losses, pos = [], 0
    for t in range(1, n_steps + 1):
        if pos >= N:
            pos = 0
        c, y = coords[pos:pos + batch], gt[pos:pos + batch]
        pos += batch
        out = model_forward(kind, params, encode(c, encB, enc_kind), net)
        out_full, mb = out, None
        if mask is not None:                      # src/train.py:172-177
            mb = mask[pos - batch:pos]
            out, y = out[mb], y[mb]
        if loss == "HDR":
            val, g, _ = loss_hdr(out.detach(), y, c, **loss_opts)
        elif loss == "LSL":
            val, g = loss_logspace(out.detach(), y, loss_opts["eps"])
        else:
            val, g = LOSS_TRAIN[loss](out.detach(), y)
        live = [p for k, p in params.items() if k not in frozen]
        if tv is not None and mb is not None and out_full.shape[0] == tv[0] * tv[1]:     # per-coil TV, src/train.py:173-174
            val_tv, g_full = loss_tv(out_full.detach(), tv[0], tv[1], tv[2] if len(tv) > 2 else 1e-4)
            g_full = g_full.clone()
            g_full[mb] += g
            val = val + val_tv
            out, g = out_full, g_full
        grads = torch.autograd.grad(out, live, grad_outputs=g)
        with torch.no_grad():
            for (k, p), gr in zip([(k, p) for k, p in params.items() if k not in frozen], grads):
                m, v = state[k]
                pr = torch.view_as_real(p) if p.is_complex() else p
                gr = torch.view_as_real(gr.contiguous()) if gr.is_complex() else gr
                adam_step(pr, gr, m, v, t, lr, betas[0], betas[1])
        losses.append(float(val))
    return losses, OrderedDict((k, v.detach()) for k, v in params.items())
