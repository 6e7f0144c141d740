"""nn.Module faces of the fused models: same constructor arguments, attribute tree and state_dict keys as
the reference classes (src/models/networks.py), parameters are views into one flat fp32 buffer that the
CUDA kernels read and update in place.

Two ways to run them:
  * unfused, exactly like the reference loop: ``out = model(x); loss.backward(); optim.step()`` --
    forward/backward are one C-ABI call each (torch.autograd.Function), any PyTorch loss / optimiser works;
  * fused: ``FusedTrainer`` (trainer.py) runs encoder + model + loss + backward + Adam as one launch sequence.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import init as pinit
from .engine import ChainEngine, Plan


class _Leaf(nn.Module):
    """Holds the parameters of one nn.Linear-shaped tensor pair under the reference's attribute names."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)


class _Wrap(nn.Module):
    def __init__(self, name: str, leaf: nn.Module):
        super().__init__()
        setattr(self, name, leaf)


class _ChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, dist, *params):
        need_grad = any(ctx.needs_input_grad[3:])      # grad mode is off inside Function.forward; ask autograd instead
        out = module._engine_forward(x, train=need_grad, dist=dist)
        ctx.module = module
        ctx.dist = dist
        ctx.bs = x.shape[0]
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        m = ctx.module
        dz = dout
        if m._act_in_kernel():
            pass          # wide chains / WIRE2D: the backward entry kernel applies the output activation's derivative itself
        elif m._last_act == "tanh":
            dz = dout * (1 - out * out)
        elif m._last_act == "sigmoid":
            dz = dout * out * (1 - out)
        grads = m._engine_backward(dz.contiguous(), ctx.bs, dist=ctx.dist)
        dead = m._dead_param_flags()
        return (None, None, None, *[g.clone() if (need and not d) else None
                                    for g, need, d in zip(grads, ctx.needs_input_grad[3:], dead)])


class FusedChain(nn.Module):
    """Base of SIREN / FFN.  ``params`` is config['net'] exactly as the reference receives it."""

    MODEL = None            # "SIREN" | "FFN"
    _PREFIX = None

    def __init__(self, params: dict):
        super().__init__()
        self.net = dict(params)
        named = self._init_tensors()                            # reference init, reference RNG order
        n = sum(t.numel() * (2 if t.is_complex() else 1) for _, t in named)
        flat = torch.empty(n, dtype=torch.float32)
        self._slices = []
        off = 0
        for name, t in named:
            r = torch.view_as_real(t).reshape(-1) if t.is_complex() else t.reshape(-1)
            flat[off:off + r.numel()] = r
            self._slices.append((name, off, tuple(t.shape), t.is_complex()))
            off += r.numel()
        self._flat = flat
        self._last_act = "sigmoid" if self.MODEL == "FFN" else ("tanh" if self.net.get("last_tanh", False) else "linear")
        if self.MODEL == "SIREN" and not self.net.get("last_tanh", False) and not self.net.get("network_last_linear", True):
            self._last_act = "sin"
        self._build_tree()
        self._engines = {}
        self._state = None
        self._param_epoch = 0
        self._max_batch = 0

    # ---- module tree with the reference's key names -------------------------------------------------
    def _init_tensors(self):
        return pinit.chain_tensors(self.MODEL, self.net)

    def _views(self, flat):
        out = []
        for _, off, shape, is_complex in self._slices:
            n = 1
            for d in shape:
                n *= d
            if is_complex:      # interleaved (re, im) float pairs == complex64 storage
                out.append(torch.view_as_complex(flat[off:off + 2 * n].view(*shape, 2)))
            else:
                out.append(flat[off:off + n].view(shape))
        return out

    def _build_tree(self):
        raise NotImplementedError

    def _leaves(self) -> List[_Leaf]:
        raise NotImplementedError

    def _params_in_order(self):
        """Parameters in flat-buffer (= state_dict) order."""
        return [p for leaf in self._leaves() for p in (leaf.weight, leaf.bias)]

    def _rebind(self, flat: torch.Tensor):
        self._flat = flat
        for p, v in zip(self._params_in_order(), self._views(flat)):
            p.data = v
            p.grad = None
        self._engines = {}
        self._state = None

    def _apply(self, fn, recurse=True):
        # keep every parameter a view of ONE flat buffer across .to()/.cuda()/.float()
        self._rebind(fn(self._flat))
        return self

    def load_state_dict(self, state_dict, strict=True, assign=False):
        res = super().load_state_dict(state_dict, strict=strict)      # copies into the views in place
        self._param_epoch += 1
        return res

    # ---- engines ------------------------------------------------------------------------------------
    def _shared_state(self, max_batch: int):
        dev = self._flat.device
        if dev.type != "cuda":
            raise L.InrError("the fused models run on CUDA only (no CPU fallback); call model.to('cuda') first")
        if self._state is None:
            z = lambda: torch.zeros_like(self._flat)
            self._state = {"grads": z(), "exp_avg": z(), "exp_avg_sq": z(),
                           "step": torch.zeros(1, dtype=torch.int32, device=dev)}
        return self._state

    def engine(self, encoder: Optional[dict], max_batch: int) -> ChainEngine:
        """Engine for this model with the given input mode (None/'none': dense [bs,in] input; gauss: in-kernel
        encoding of coords).  All engines of a module share its parameters, gradients and Adam state."""
        key = encoder.get("embedding") if (encoder and encoder.get("embedding") in ("gauss", "LogF")) else "none"
        st = self._shared_state(max_batch)
        eng = self._engines.get(key)
        if eng is None or eng.max_batch < max_batch:
            plan = Plan(self.MODEL, self.net, encoder if key != "none" else {"embedding": "none"})
            if self.MODEL in ("WIRE", "WIRE2D"):
                assert key == "none", "WIRE / WIRE2D take raw coordinates"
            eng = ChainEngine(plan, max_batch=max(max_batch, 128), device=self._flat.device,
                              shared={"params": self._flat, **st})
            eng.packed_epoch = -1
            self._engines[key] = eng
        if eng.packed_epoch != self._epoch_token():
            eng.pack()
            eng.packed_epoch = self._epoch_token()
        return eng

    def _epoch_token(self):
        # In-place edits through torch bump a version counter: writes to the flat buffer bump _flat._version, but
        # an optimiser step (torch.optim.Adam: p.addcdiv_) or any edit through a Parameter bumps THAT Parameter's
        # counter only -- after _rebind's `p.data = view` every Parameter counts on its own.  Version counters only
        # grow, so their sum changes whenever any of them does.  Fused steps edit through raw pointers and bump
        # _param_epoch instead.
        return (self._param_epoch, self._flat._version, sum(p._version for p in self._params_in_order()))

    def mark_params_updated_by_kernel(self, eng: ChainEngine):
        self._param_epoch += 1
        eng.packed_epoch = self._epoch_token()      # the fused optimiser re-packed this engine's copies itself

    def _act_in_kernel(self):
        eng = next(iter(self._engines.values()))
        # wide chains, and WIRE2D's complex tanh tail (the derivative needs the imaginary part of the final linear as well)
        return bool(getattr(eng.plan, "wide", False)) or eng.plan.model == "WIRE2D"

    def _dead_param_flags(self):
        """True for parameters no gradient ever reaches (reference autograd leaves their .grad None)."""
        eng = next(iter(self._engines.values()))
        return [frozen for (_, frozen) in eng.plan.tensor_flags]

    def _engine_forward(self, x: torch.Tensor, train: bool, dist=None) -> torch.Tensor:
        eng = self.engine(None, x.shape[0])
        return eng.forward(x, train=train, dist=dist)

    def _engine_backward(self, dz: torch.Tensor, bs: int, dist=None):
        eng = self.engine(None, bs)
        eng.backward(dz, dist=dist)
        return self._views(eng.grads)

    def forward(self, x):
        return _ChainFn.apply(x, self, None, *self._params_in_order())


class SIREN(FusedChain):
    """reference src/models/networks.py:99-124 -- keys model.<i>.linear.{weight,bias}."""
    MODEL = "SIREN"

    def _build_tree(self):
        v = self._views(self._flat)
        layers = [_Wrap("linear", _Leaf(v[2 * i], v[2 * i + 1])) for i in range(len(v) // 2)]
        self.model = nn.Sequential(*layers)

    def _leaves(self):
        return [m.linear for m in self.model]


class FFN(FusedChain):
    """reference src/models/networks.py:48-69 -- keys model.<2i>.{weight,bias} (activations sit at odd indices)."""
    MODEL = "FFN"

    def _build_tree(self):
        v = self._views(self._flat)
        mods = []
        for i in range(len(v) // 2):
            mods.append(_Leaf(v[2 * i], v[2 * i + 1]))
            mods.append(nn.Identity())      # placeholder for ReLU / Sigmoid: keeps the reference's index spacing
        self.model = nn.Sequential(*mods)

    def _leaves(self):
        return [m for m in self.model if isinstance(m, _Leaf)]


class _GaborHolder(nn.Module):
    """One ComplexGaborLayer's parameters under the reference names: omega_0, scale_0 (frozen), linear.{weight,bias}."""

    def __init__(self, omega, scale, weight, bias):
        super().__init__()
        self.omega_0 = nn.Parameter(omega, requires_grad=False)
        self.scale_0 = nn.Parameter(scale, requires_grad=False)
        self.linear = _Leaf(weight, bias)


class WIRE(FusedChain):
    """reference src/models/networks.py:206-260 -- keys net.<i>.{omega_0,scale_0,linear.weight,linear.bias},
    net.<depth+1>.{weight,bias}; complex64 hidden / final tensors are views of interleaved (re, im) floats of the flat
    buffer.  forward(coords [bs,3]) returns the real part of the final complex linear."""
    MODEL = "WIRE"

    def _init_tensors(self):
        return pinit.wire_tensors(self.net)

    def _build_tree(self):
        v = self._views(self._flat)
        depth = self.net["network_depth"]
        mods = [_GaborHolder(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) for i in range(depth + 1)]
        mods.append(_Leaf(v[4 * (depth + 1)], v[4 * (depth + 1) + 1]))
        self.net_modules = mods
        self.net_tree = nn.Sequential(*mods)

    def __setattr__(self, name, value):
        # the reference keeps both the config dict (constructor arg) and the nn.Sequential under the name `net`;
        # state_dict keys need the Sequential registered as "net"
        super().__setattr__(name, value)

    def _params_in_order(self):
        out = []
        for m in self.net_modules[:-1]:
            out += [m.omega_0, m.scale_0, m.linear.weight, m.linear.bias]
        out += [self.net_modules[-1].weight, self.net_modules[-1].bias]
        return out

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return type(sd)((k.replace("net_tree.", "net.", 1), v) for k, v in sd.items())

    def load_state_dict(self, state_dict, strict=True, assign=False):
        renamed = {k.replace("net.", "net_tree.", 1) if k.startswith("net.") else k: v for k, v in state_dict.items()}
        res = nn.Module.load_state_dict(self, renamed, strict=strict)
        self._param_epoch += 1
        return res


class _Gabor2DHolder(nn.Module):
    """One ComplexGaborLayer2D's parameters (wire2d.py:35-46): omega_0, scale_0 (frozen), linear.*, scale_orth.*."""

    def __init__(self, omega, scale, weight, bias, oweight, obias):
        super().__init__()
        self.omega_0 = nn.Parameter(omega, requires_grad=False)
        self.scale_0 = nn.Parameter(scale, requires_grad=False)
        self.linear = _Leaf(weight, bias)
        self.scale_orth = _Leaf(oweight, obias)


class WIRE2D(WIRE):
    """reference src/models/wire2d.py:63-118 -- keys net.<i>.{omega_0,scale_0,linear.*,scale_orth.*}, net.<depth+1>.*;
    hidden width = network_width (not reduced).  forward(coords [bs,3]) returns the real part of the final linear."""
    MODEL = "WIRE2D"

    def _init_tensors(self):
        return pinit.wire2d_tensors(self.net)

    def _build_tree(self):
        v = self._views(self._flat)
        depth = self.net["network_depth"]
        mods = [_Gabor2DHolder(*v[6 * i:6 * i + 6]) for i in range(depth + 1)]
        mods.append(_Leaf(v[6 * (depth + 1)], v[6 * (depth + 1) + 1]))
        self.net_modules = mods
        self.net_tree = nn.Sequential(*mods)

    def _params_in_order(self):
        out = []
        for m in self.net_modules[:-1]:
            out += [m.omega_0, m.scale_0, m.linear.weight, m.linear.bias, m.scale_orth.weight, m.scale_orth.bias]
        out += [self.net_modules[-1].weight, self.net_modules[-1].bias]
        return out


class FourierNet(FusedChain):
    """reference src/models/mfn.py:61-94 -- keys linear.<i>.{weight,bias}, output_linear.{weight,bias},
    filters.<i>.linear.{weight,bias}.  forward(x [bs, in]) -> [bs, out]."""
    MODEL = "Fourier"
    N_OUT = None

    def __init__(self, params, out_size=1.0, input_scale=2.0, weight_scale=1.0, bias=True, output_act=False):
        if output_act or not bias:
            raise L.InrError("output_act / bias=False variants of the MFNs are not built")
        self._scales = (input_scale, weight_scale)
        super().__init__(params)

    def _init_tensors(self):
        return pinit.mfn_tensors(self.MODEL, self.net, *self._scales)

    def _build_tree(self):
        v = self._views(self._flat)
        L_ = self.net["network_depth"]
        self.linear = nn.ModuleList([self._lin_holder(v[2 * i], v[2 * i + 1]) for i in range(L_)])
        nh = self._n_head_tensors()
        o = 2 * L_
        if nh == 1:
            self.output_linear = _Leaf(v[o], v[o + 1])
        else:
            self.output_linear = nn.ModuleList([_Leaf(v[o + 2 * k], v[o + 2 * k + 1]) for k in range(nh)])
        o += 2 * nh
        self.filters = nn.ModuleList([_Wrap("linear", _Leaf(v[o + 2 * i], v[o + 2 * i + 1])) for i in range(L_ + 1)])

    def _lin_holder(self, w, b):
        return _Leaf(w, b)

    def _n_head_tensors(self):
        return 1

    def _leaves(self):
        lins = [m if isinstance(m, _Leaf) else m.linear for m in self.linear]
        heads = [self.output_linear] if isinstance(self.output_linear, _Leaf) else list(self.output_linear)
        return lins + heads + [f.linear for f in self.filters]

    def forward(self, x, dist_to_center=None):
        return _ChainFn.apply(x, self, None, *self._params_in_order())


class _GaborFilterHolder(nn.Module):
    """Parameters of one reference GaborLayer (mfn.py:102-114) under its names: mu, gamma, linear.{weight,bias}."""

    def __init__(self, mu, gamma, weight, bias):
        super().__init__()
        self.mu = nn.Parameter(mu)
        self.gamma = nn.Parameter(gamma)
        self.linear = _Leaf(weight, bias)


class GaborNet(FourierNet):
    """reference src/models/mfn.py:133-162 -- keys linear.<i>.*, output_linear.*, filters.<i>.{mu, gamma, linear.weight,
    linear.bias}.  forward(x [bs, in]) -> [bs, out]."""
    MODEL = "Gabor"

    def __init__(self, params, input_scale=2, weight_scale=1.0, alpha=6.0, beta=1.0, bias=True, output_act=False):
        if output_act or not bias:
            raise L.InrError("output_act / bias=False variants of the MFNs are not built")
        self._scales = (float(input_scale), float(weight_scale), float(alpha), float(beta))
        FusedChain.__init__(self, params)

    def _build_tree(self):
        v = self._views(self._flat)
        L_ = self.net["network_depth"]
        self.linear = nn.ModuleList([_Leaf(v[2 * i], v[2 * i + 1]) for i in range(L_)])
        o = 2 * L_
        self.output_linear = _Leaf(v[o], v[o + 1])
        o += 2
        self.filters = nn.ModuleList([_GaborFilterHolder(v[o + 4 * i], v[o + 4 * i + 1], v[o + 4 * i + 2], v[o + 4 * i + 3])
                                      for i in range(L_ + 1)])

    def _params_in_order(self):
        out = []
        for m in self.linear:
            out += [m.weight, m.bias]
        out += [self.output_linear.weight, self.output_linear.bias]
        for f in self.filters:
            out += [f.mu, f.gamma, f.linear.weight, f.linear.bias]
        return out


class KGaborNet(GaborNet):
    """reference src/models/mfn.py:164-204 -- forward(x, dist_to_center); the reference never enables
    with_dist_filtering, so dist_to_center is accepted and ignored exactly as there (:184-198)."""
    MODEL = "KGabor"

    def forward(self, x, dist_to_center=None):
        return GaborNet.forward(self, x)


class MultiscaleKFourier(FourierNet):
    """reference src/models/mfn.py:206-267 -- heads output_linear.<i> at `output_layers`; forward(coords=x) returns the
    LIST of head outputs in stage order.  Parameters of dead stages / unused heads never receive gradients."""
    MODEL = "MultiscaleFourier"

    def __init__(self, params, weight_scale=1.0, bias=True, output_act=False, centered=True, output_layers=(1, 3, 5, 7),
                 reuse_filters=False):
        params = dict(params)
        params["output_layers"] = list(output_layers) if output_layers is not None else list(range(1, params["network_depth"] + 1))
        self.output_layers = params["output_layers"]
        self.stop_after = None
        FourierNet.__init__(self, params, input_scale=2.0, weight_scale=weight_scale, bias=bias, output_act=output_act)

    def _n_head_tensors(self):
        return self.net["network_depth"] + 1

    def forward(self, coords, **args):
        out = _ChainFn.apply(coords, self, None, *self._params_in_order())
        f = self.net["network_output_size"]
        return [out[:, f * k:f * (k + 1)] for k in range(len(self.output_layers))]


class _BoundedHolder(nn.Module):
    def __init__(self, weight, bias, bounds):
        super().__init__()
        self.linear = _Leaf(weight, bias)
        self.bounds = bounds


class MultiscaleBoundedFourier(MultiscaleKFourier):
    """reference src/models/mfn.py:288-356 -- BoundedLinear layers (keys linear.<i>.linear.{weight,bias}) zero the rows whose
    distance to the k-space centre lies outside `boundaries[i]` before linear i (1-D dist_to_center: whole rows)."""
    MODEL = "BoundedFourier"

    def __init__(self, params, weight_scale=1.0, bias=True, output_act=False, centered=True, output_layers=(1, 3, 5, 7),
                 reuse_filters=False, boundaries=None):
        params = dict(params)
        if boundaries is None or len(boundaries) < params["network_depth"]:
            raise L.InrError("MultiscaleBoundedFourier needs one (lo, hi) boundary per linear layer")
        params["boundaries"] = [tuple(float(v) for v in b) for b in boundaries]
        MultiscaleKFourier.__init__(self, params, weight_scale, bias, output_act, centered, output_layers, reuse_filters)

    def _lin_holder(self, w, b):
        return _BoundedHolder(w, b, None)

    def forward(self, coords, dist_to_center=None):
        if dist_to_center is None:
            raise L.InrError("MultiscaleBoundedFourier.forward needs dist_to_center")
        d = dist_to_center
        if d.dim() != 1:
            raise L.InrError("only the 1-D dist_to_center convention (whole rows zeroed) is built; the per-coil [HW,1] quirk is not")
        out = _ChainFn.apply(coords, self, d.to(coords.device, torch.float32).contiguous(), *self._params_in_order())
        f = self.net["network_output_size"]
        return [out[:, f * k:f * (k + 1)] for k in range(len(self.output_layers))]


class Positional_Encoder:
    """reference src/models/networks.py:7-35 (same constructor, .B, .embedding)."""

    def __init__(self, params: dict, device):
        self.device = device
        self.params = dict(params)
        self.embedding_type = params["embedding"]
        self.B = pinit.encoder_matrix(params)
        if params["embedding"] == "LogF":
            steps = int(params["embedding_size"] / (2 * params["coordinates_size"]))
            self.B = (2.0 ** torch.linspace(0.0, params["scale"], steps=steps)).reshape(-1, 1)
        elif params["embedding"] not in ("gauss", "none"):
            raise NotImplementedError
        if self.B is not None:
            self.B = self.B.to(device)

    def embedding(self, x):
        two_pi = 6.283185307179586
        if self.embedding_type == "LogF":
            parts = []
            for c in range(3):
                a = (two_pi * x[:, c:c + 1]) @ self.B.T
                parts += [torch.sin(a), torch.cos(a)]
            return torch.cat(parts, dim=-1)
        if self.B is not None:
            a = (two_pi * x) @ self.B.t()
            return torch.cat([torch.sin(a), torch.cos(a)], dim=-1)
        return x
