"""Fused training path: device-resident grid-order batcher + one launch sequence per step.

Replaces the body of the reference loops (src/train.py:158-192) when model, loss and regulariser are all
fusable; keeps the reference's semantics: shuffle=False grid-order batches (src/models/utils.py:84-90), a
short last batch, row mask (src/train.py:172-177), 0.5 weight on L2/L1/MSLE (:182), Adam + per-epoch
LambdaLR (:76,:153,:251), regulariser added to the loss (:185-187)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .modules import FusedChain, Positional_Encoder

FUSABLE_LOSSES = ("L2", "L1", "MSLE", "tanh", "HDR", "LSL")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam-compatible face (param_groups for LambdaLR, state_dict layout) over the engine's flat
    Adam buffers.  ``step()`` serves the unfused path (grads produced by autograd); the fused path updates the
    parameters inside ``FusedTrainer.step`` and ``step()`` is then a no-op for that batch."""

    def __init__(self, model: FusedChain, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.model = model
        super().__init__(list(model.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._fused_done = False
        self.reg_l1 = 0.0
        self.reg_l2 = 0.0

    def hyper(self):
        g = self.param_groups[0]
        return [g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], self.reg_l1, self.reg_l2, 0.0]

    def sync_hyper(self, eng):
        h = self.hyper()
        if getattr(eng, "_hyper_host", None) != h:
            eng.hyper.copy_(torch.tensor(h, dtype=torch.float32))
            eng._hyper_host = h

    @torch.no_grad()
    def step(self, closure=None):
        if self._fused_done:
            self._fused_done = False
            return None
        m = self.model
        eng = m.engine(None, 128)
        for v, p in zip(m._views(eng.grads), self.param_groups[0]["params"]):
            if p.grad is not None:
                v.copy_(p.grad)
            else:
                v.zero_()
        self.sync_hyper(eng)
        eng.adam_step()
        m.mark_params_updated_by_kernel(eng)
        return None

    def state_dict(self):
        sd = super().state_dict()
        st = self.model._state
        if st is not None:
            step = st["step"].detach().float().cpu().reshape(())
            for i, (ea, es) in enumerate(zip(self.model._views(st["exp_avg"]), self.model._views(st["exp_avg_sq"]))):
                sd["state"][i] = {"step": step.clone(), "exp_avg": ea.detach().clone(), "exp_avg_sq": es.detach().clone()}
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        for g, src in zip(self.param_groups, groups):
            for k, v in src.items():
                if k != "params":
                    g[k] = v
        st = self.model._shared_state(128)
        for i, (ea, es) in enumerate(zip(self.model._views(st["exp_avg"]), self.model._views(st["exp_avg_sq"]))):
            s = state_dict["state"].get(i)
            if s is not None:
                ea.copy_(s["exp_avg"].to(ea.device))
                es.copy_(s["exp_avg_sq"].to(es.device))
                st["step"].fill_(int(s["step"]))


class FusedTrainer:
    """Owns the resident arrays of one slice and walks them in grid order, one fused step per batch."""

    def __init__(self, model: FusedChain, encoder: Positional_Encoder, optim: FusedAdam, loss: str, batch_size: int,
                 coords: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None, loss_opts: Optional[dict] = None,
                 use_graph: bool = True, tv: Optional[tuple] = None):
        """tv = (H, W[, weight]): per-coil batches (batch_size == H * W) with the total-variation term of
        src/train.py:173-174 added on every batch (the reference only reaches it with an undersampling mask)."""
        if loss not in FUSABLE_LOSSES:
            raise L.InrError(f"loss '{loss}' is not fused; use the unfused model(x)/backward path")
        wire = getattr(model, "MODEL", None) in ("WIRE", "WIRE2D")
        if wire and encoder.embedding_type != "none":
            raise L.InrError("WIRE is fitted on raw coordinates (encoder.embedding: none)")
        if not wire and encoder.embedding_type not in ("gauss",):
            raise L.InrError("the fused step needs the gauss encoder (dense inputs go through the unfused path)")
        dev = model._flat.device
        self.model, self.encoder, self.optim = model, encoder, optim
        self.loss, self.loss_opts, self.bs = loss, dict(loss_opts or {}), int(batch_size)
        if tv is not None:
            if int(tv[0]) * int(tv[1]) != int(batch_size):
                raise L.InrError("the TV term needs per-coil batches: batch_size == H * W")
            self.loss_opts["tv"] = tuple(tv)
        self.coords = coords.to(dev, torch.float32).contiguous()
        self.gt = gt.to(dev, torch.float32).contiguous()
        self.mask = None if mask is None else mask.to(dev).to(torch.uint8).contiguous()
        self.n = self.coords.shape[0]
        self._enc_params = None if wire else encoder.params
        self.eng = model.engine(self._enc_params, self.bs)
        self.eng.set_encoder(encoder.B)
        self._out = (torch.empty(self.bs, self.eng.plan.out_cols, dtype=torch.float32, device=dev)
                     if "tv" in self.loss_opts else None)
        self.use_graph = use_graph
        self._graphs = {}
        self.pos = 0
        self.eng.cursor.zero_()

    @property
    def steps_per_epoch(self):
        return (self.n + self.bs - 1) // self.bs

    def _launch(self, bs):
        self.eng.train_step(self.loss, self.coords, self.gt, bs, mask=self.mask, loss_opts=self.loss_opts, use_cursor=True,
                            out=self._out)

    def step(self) -> torch.Tensor:
        """One batch; returns the device scalar holding its loss (no host sync)."""
        if self.pos >= self.n:
            self.pos = 0
            self.eng.cursor.zero_()
        bs = min(self.bs, self.n - self.pos)
        self.optim.sync_hyper(self.eng)
        # engines of this module share the parameters: make sure this one's fp16 copies are current
        self.model.engine(self._enc_params, self.bs)
        g = self._graphs.get(bs) if self.use_graph else None
        if g is not None:
            g.replay()
        else:
            self._launch(bs)                          # eager (first step of this batch size, or graphs off)
            if self.use_graph:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):             # capture records the same four launches; nothing executes here
                    self._launch(bs)
                self._graphs[bs] = g
        self.pos += bs
        self.model.mark_params_updated_by_kernel(self.eng)
        self.optim._fused_done = False
        return self.eng.loss_out

    def epoch(self):
        """All batches of one pass in grid order; returns the summed loss as a device tensor."""
        total = torch.zeros((), device=self.coords.device)
        for _ in range(self.steps_per_epoch):
            total += self.step().squeeze()
        return total

    @torch.no_grad()
    def predict(self, coords: Optional[torch.Tensor] = None, chunk: Optional[int] = None) -> torch.Tensor:
        """Full-grid inference (validation, src/train.py:199-220) through the fused forward, chunked."""
        coords = self.coords if coords is None else coords.to(self.coords.device, torch.float32).contiguous()
        chunk = chunk or self.bs
        eng = self.model.engine(self._enc_params, chunk)
        eng.set_encoder(self.encoder.B)
        outs = [eng.forward(coords[i:i + chunk], train=False) for i in range(0, coords.shape[0], chunk)]
        return torch.cat(outs)
