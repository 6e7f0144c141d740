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
    def forward(ctx, x, module, *params):
        need_grad = any(ctx.needs_input_grad[2:])      # grad mode is off inside Function.forward; ask autograd instead
        out = module._engine_forward(x, train=need_grad)
        ctx.module = module
        ctx.bs = x.shape[0]
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        m = ctx.module
        dz = dout
        if m._last_act == "tanh":
            dz = dout * (1 - out * out)
        elif m._last_act == "sigmoid":
            dz = dout * out * (1 - out)
        grads = m._engine_backward(dz.contiguous(), ctx.bs)
        return (None, None, *[g.clone() for g in grads])


class FusedChain(nn.Module):
    """Base of SIREN / FFN.  ``params`` is config['net'] exactly as the reference receives it."""

    MODEL = None            # "SIREN" | "FFN"
    _PREFIX = None

    def __init__(self, params: dict):
        super().__init__()
        self.net = dict(params)
        named = pinit.chain_tensors(self.MODEL, self.net)       # reference init, reference RNG order
        n = sum(t.numel() for _, t in named)
        flat = torch.empty(n, dtype=torch.float32)
        self._slices = []
        off = 0
        for name, t in named:
            flat[off:off + t.numel()] = t.reshape(-1)
            self._slices.append((name, off, tuple(t.shape)))
            off += t.numel()
        self._flat = flat
        self._last_act = "sigmoid" if self.MODEL == "FFN" else ("tanh" if self.net.get("last_tanh", False) else "linear")
        self._build_tree()
        self._engines = {}
        self._state = None
        self._param_epoch = 0
        self._max_batch = 0

    # ---- module tree with the reference's key names -------------------------------------------------
    def _views(self, flat):
        return [flat[off:off + int(torch.tensor(shape).prod())].view(shape) for _, off, shape in self._slices]

    def _build_tree(self):
        raise NotImplementedError

    def _leaves(self) -> List[_Leaf]:
        raise NotImplementedError

    def _rebind(self, flat: torch.Tensor):
        self._flat = flat
        views = self._views(flat)
        for i, leaf in enumerate(self._leaves()):
            leaf.weight.data = views[2 * i]
            leaf.bias.data = views[2 * i + 1]
            leaf.weight.grad = None
            leaf.bias.grad = None
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
        key = "gauss" if (encoder and encoder.get("embedding") == "gauss") else "none"
        st = self._shared_state(max_batch)
        eng = self._engines.get(key)
        if eng is None or eng.max_batch < max_batch:
            plan = Plan(self.MODEL, self.net, encoder if key == "gauss" else {"embedding": "none"})
            eng = ChainEngine(plan, max_batch=max(max_batch, 128), device=self._flat.device,
                              shared={"params": self._flat, **st})
            eng.packed_epoch = -1
            self._engines[key] = eng
        if eng.packed_epoch != self._epoch_token():
            eng.pack()
            eng.packed_epoch = self._epoch_token()
        return eng

    def _epoch_token(self):
        # in-place edits through torch (optimiser steps, load_state_dict, manual writes) bump _version;
        # fused steps edit through raw pointers and bump _param_epoch instead
        return (self._param_epoch, self._flat._version)

    def mark_params_updated_by_kernel(self, eng: ChainEngine):
        self._param_epoch += 1
        eng.packed_epoch = self._epoch_token()      # the fused optimiser re-packed this engine's copies itself

    def _engine_forward(self, x: torch.Tensor, train: bool) -> torch.Tensor:
        eng = self.engine(None, x.shape[0])
        return eng.forward(x, train=train)

    def _engine_backward(self, dz: torch.Tensor, bs: int):
        eng = self.engine(None, bs)
        eng.backward(dz)
        return self._views(eng.grads)

    def forward(self, x):
        params = [p for leaf in self._leaves() for p in (leaf.weight, leaf.bias)]
        return _ChainFn.apply(x, self, *params)


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
