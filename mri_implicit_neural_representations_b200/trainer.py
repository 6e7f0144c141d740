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


def dp_batch_table(coords: torch.Tensor, mask: Optional[torch.Tensor], batch_size: int, world: int, loss: str,
                   loss_opts: Optional[dict]) -> torch.Tensor:
    """Loss normalisers of every GLOBAL grid-order batch of a slice, [n_batches, 2] fp32 (inr_loss_desc.dp_norm):
    column 0 = rows of the batch that enter the loss / world, column 1 = the HDR filter mean
    mean_i (1 - exp(-(k_i1^2 + k_i2^2) / (2 sigma^2)))^2 over ALL rows of the batch (reference src/metrics/losses.py:241-259).
    Both depend on the inputs only, so every rank computes the same table without communication."""
    n = coords.shape[0]
    nb = (n + batch_size - 1) // batch_size
    pad = nb * batch_size - n
    ones = torch.ones(n, dtype=torch.float32, device=coords.device) if mask is None else (mask.reshape(-1) != 0).float()
    cnt = torch.nn.functional.pad(ones, (0, pad)).view(nb, batch_size).sum(1)
    fmean = torch.zeros(nb, dtype=torch.float32, device=coords.device)
    if loss == "HDR":
        sigma = float((loss_opts or {}).get("hdr_ff_sigma", 1.0))
        f = torch.exp(-(coords[:, 1] ** 2 + coords[:, 2] ** 2) / (2.0 * sigma * sigma))
        w = (1.0 - f) ** 2
        rows = torch.nn.functional.pad(torch.ones_like(w), (0, pad)).view(nb, batch_size).sum(1)
        fmean = torch.nn.functional.pad(w, (0, pad)).view(nb, batch_size).sum(1) / rows
    return torch.stack([cnt / float(world), fmean], dim=1).contiguous()


class DataParallel:
    """Coordinate data-parallel context of one fit (nothing like it exists in the reference, whose loop is
    single-device): one process per GPU, replicated weights, every rank takes a contiguous share of each global grid-order
    batch, gradients are exchanged once per step and every rank applies the same Adam update.

    The exchange is the optimiser-fused peer gather over NVLink (`parallel.PeerGradExchange`) when symmetric memory can
    be set up on every rank, else one NCCL all-reduce(avg) of the flat fp32 gradient buffer."""

    def __init__(self, group=None, exchange: Optional[str] = None):
        import os
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise L.InrError("DataParallel needs an initialised torch.distributed process group (launch with torchrun)")
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.exchange = exchange or os.environ.get("INR_DP_EXCHANGE", "peer")
        self.peer = None

    def setup(self, n_params: int, device):
        """Collective: all ranks agree on the exchange mechanism."""
        dist = self.dist
        if self.world == 1:
            return
        ok = 0
        if self.exchange == "peer" and dist.get_backend(self.group) == "nccl":
            try:
                from .parallel import PeerGradExchange
                self.peer = PeerGradExchange(n_params, device, self.group)
                ok = 1
            except Exception:
                self.peer = None
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag) == 0:
            self.peer = None

    def describe(self) -> str:
        return "peer gather fused into the optimiser kernel (NVLink symmetric memory)" if self.peer is not None else "NCCL all-reduce(avg)"


class FusedTrainer:
    """Owns the resident arrays of one slice and walks them in grid order, one fused step per batch.

    With ``dp`` (a `DataParallel`) the SAME global batches are walked by all ranks together: rank r processes its
    contiguous share of every global batch (`parallel.shard_rows`), the loss means run over the global batch
    (`dp_batch_table`), gradients are averaged over ranks and every rank applies the same update -- so a world-size-R
    fit follows the single-process trajectory up to summation order."""

    def __init__(self, model: FusedChain, encoder: Positional_Encoder, optim: FusedAdam, loss: str, batch_size: int,
                 coords: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None, loss_opts: Optional[dict] = None,
                 use_graph: bool = True, tv: Optional[tuple] = None, dp: Optional[DataParallel] = None,
                 dist: Optional[torch.Tensor] = None, consistency: Optional[tuple] = None):
        """tv = (H, W[, weight]): per-coil batches (batch_size == H * W) with the total-variation term of
        src/train.py:173-174 added on every batch (the reference only reaches it with an undersampling mask).
        dist / consistency = (bounds, weight): the multi-scale loop of src/train_kspace_multiscale.py:164-192 -- the model is
        called with dist_to_center, every head gets `loss` on the full target plus weight * ConsistencyLoss(bounds)."""
        if loss not in FUSABLE_LOSSES:
            raise L.InrError(f"loss '{loss}' is not fused; use the unfused model(x)/backward path")
        wire = getattr(model, "MODEL", None) in ("WIRE", "WIRE2D")
        if wire and encoder.embedding_type != "none":
            raise L.InrError("WIRE is fitted on raw coordinates (encoder.embedding: none)")
        logf_ok = encoder.embedding_type == "LogF" and getattr(model, "MODEL", None) in ("SIREN", "FFN")
        if not wire and encoder.embedding_type != "gauss" and not logf_ok:
            raise L.InrError("the fused step needs the gauss encoder, or LogF with SIREN / FFN (dense inputs go through the unfused path)")
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
        self.dist = None if dist is None else dist.to(dev, torch.float32).reshape(-1).contiguous()
        multi = getattr(model, "MODEL", None) in ("MultiscaleFourier", "BoundedFourier")
        if multi:
            if self.dist is None or self.dist.numel() != self.n:
                raise L.InrError("multi-scale models need dist_to_center for every row")
            if loss not in ("L2", "L1", "MSLE", "LSL") or tv is not None:
                raise L.InrError(f"the fused multi-scale step handles L2 / L1 / MSLE / LSL without TV, not '{loss}'"
                                 + (" + TV" if tv is not None else ""))
            if dp is not None and dp.world > 1:
                raise L.InrError("multi-scale fits are not coordinate data-parallel; place whole fits (or ring models) on GPUs")
            if consistency is not None:
                self.loss_opts["consistency"] = (list(consistency[0]), float(consistency[1]))
        self.dp = dp if (dp is not None and dp.world > 1) else None
        self.global_bs, self.n_global = self.bs, self.n
        self._full_coords = self.coords               # validation predicts the whole grid on every rank
        if self.dp is not None:
            self._shard_rows(tv)
        self._enc_params = None if wire else encoder.params
        self.eng = model.engine(self._enc_params, self.bs)
        self.eng.set_encoder(encoder.B)
        self._out = (torch.empty(self.bs, self.eng.plan.out_cols, dtype=torch.float32, device=dev)
                     if "tv" in self.loss_opts else None)
        self.use_graph = use_graph
        self._graphs = {}
        self.pos = 0
        self._gstep = 0                               # optimiser steps taken: parity of the peer exchange buffers
        self._batch = 0                               # index of the next batch inside the epoch
        self.eng.cursor.zero_()
        if self.dp is not None:
            self.dp.setup(self.eng.plan.n_params, dev)

    def _shard_rows(self, tv):
        """Data-parallel: keep this rank's share of every global batch, contiguous and in batch order, so the local walk is
        again a plain grid-order walk with batch size global_bs / world (the short last batch is split to within a row)."""
        from .parallel import shard_rows
        dp = self.dp
        if tv is not None:
            raise L.InrError("per-coil TV batches are not sharded over ranks (the TV term couples neighbouring rows); "
                             "run per-coil fits as independent fits, one per GPU")
        if self.global_bs % dp.world != 0:
            raise L.InrError(f"data-parallel fits need batch_size ({self.global_bs}) divisible by the world size ({dp.world})")
        self._dp_table = dp_batch_table(self.coords, self.mask, self.global_bs, dp.world, self.loss, self.loss_opts)
        idx, self._local_counts = [], []
        for start in range(0, self.n, self.global_bs):
            s, c = shard_rows(start, self.global_bs, self.n, dp.rank, dp.world)
            idx.append(torch.arange(s, s + c, device=self.coords.device))
            self._local_counts.append(c)
        idx = torch.cat(idx)
        self.coords = self.coords[idx].contiguous()
        self.gt = self.gt[idx].contiguous()
        self.mask = None if self.mask is None else self.mask[idx].contiguous()
        self.bs = self.global_bs // dp.world
        self.n = int(idx.numel())
        self.loss_opts["dp_norm"] = (self._dp_table, self.bs)

    @property
    def steps_per_epoch(self):
        return (self.n_global + self.global_bs - 1) // self.global_bs

    def _launch(self, bs, par=0):
        if self.dp is None:
            self.eng.train_step(self.loss, self.coords, self.gt, bs, mask=self.mask, loss_opts=self.loss_opts, use_cursor=True,
                                out=self._out, dist=self.dist)
        elif self.dp.peer is not None:
            # forward + loss + backward into this rank's peer-mapped buffer of parity `par`; the optimiser kernel waits for
            # every rank's flag, gathers all ranks' gradients over NVLink and applies Adam (no all-reduce kernel)
            self.eng.grad_step(self.loss, self.coords, self.gt, bs, mask=self.mask, loss_opts=self.loss_opts, use_cursor=True,
                               out=self._out, grads=self.dp.peer.grads(par))
            self.eng.adam_step_peers(self.dp.peer, par)
        else:
            from .parallel import allreduce_mean_
            self.eng.grad_step(self.loss, self.coords, self.gt, bs, mask=self.mask, loss_opts=self.loss_opts, use_cursor=True,
                               out=self._out)
            allreduce_mean_(self.eng.grads, group=self.dp.group)
            self.eng.adam_step()

    def step(self) -> torch.Tensor:
        """One batch; returns the device scalar holding its loss (no host sync).  Data-parallel: the scalar is this rank's
        share, normalised so that the mean over ranks is the global-batch loss (`global_loss`)."""
        if self.pos >= self.n:
            self.pos = 0
            self._batch = 0
            self.eng.cursor.zero_()
        if self.dp is None:
            bs = min(self.bs, self.n - self.pos)
        else:
            bs = self._local_counts[self._batch]
            if bs == 0:
                raise L.InrError("a rank received no rows of the last global batch; use a batch size the slice does not "
                                 "leave a remainder smaller than the world size for")
        self.optim.sync_hyper(self.eng)
        # engines of this module share the parameters: make sure this one's fp16 copies are current
        self.model.engine(self._enc_params, self.bs)
        par = self._gstep & 1 if (self.dp is not None and self.dp.peer is not None) else 0
        key = (bs, par)
        g = self._graphs.get(key) if self.use_graph else None
        if g is not None:
            g.replay()
        else:
            self._launch(bs, par)                     # eager (first step of this batch size / parity, or graphs off)
            if self.use_graph and self._gstep >= 1:   # calibration passes and kernel attributes are behind us
                torch.cuda.synchronize()
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):         # capture only records the launches; nothing executes here
                        self._launch(bs, par)
                    self._graphs[key] = g
                except Exception:
                    if self.dp is None:
                        raise
                    self.use_graph = False            # a collective that refuses capture: stay eager
                    torch.cuda.synchronize()
        self.pos += bs
        self._batch += 1
        self._gstep += 1
        self.model.mark_params_updated_by_kernel(self.eng)
        self.optim._fused_done = False
        return self.eng.loss_out

    def global_loss(self, loss_dev: torch.Tensor) -> float:
        """Host value of a step's loss; data-parallel: mean over ranks (one tiny all-reduce -- call it at log_iter only)."""
        if self.dp is None:
            return float(loss_dev)
        t = loss_dev.detach().clone().reshape(1)
        self.dp.dist.all_reduce(t, op=self.dp.dist.ReduceOp.SUM, group=self.dp.group)
        return float(t) / self.dp.world

    def param_checksum(self) -> float:
        """Sum of all parameters as fp64 (data-parallel replicas must agree bit for bit)."""
        return float(self.model._flat.double().sum())

    def epoch(self):
        """All batches of one pass in grid order; returns the summed loss as a device tensor."""
        total = torch.zeros((), device=self.coords.device)
        for _ in range(self.steps_per_epoch):
            total += self.step().squeeze()
        return total

    @torch.no_grad()
    def predict(self, coords: Optional[torch.Tensor] = None, chunk: Optional[int] = None) -> torch.Tensor:
        """Full-grid inference (validation, src/train.py:199-220) through the fused forward, chunked."""
        own = coords is None
        coords = self._full_coords if own else coords.to(self.coords.device, torch.float32).contiguous()
        chunk = chunk or self.global_bs
        eng = self.model.engine(self._enc_params, chunk)
        eng.set_encoder(self.encoder.B)
        dist = None
        if self.dist is not None:        # multi-scale models: every head of every row, [N, n_heads * out]
            dist = self.dist if own else torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
        outs = [eng.forward(coords[i:i + chunk], train=False, dist=None if dist is None else dist[i:i + chunk])
                for i in range(0, coords.shape[0], chunk)]
        return torch.cat(outs)


class HostFedStepper:
    """Fused steps fed from HOST batches -- the reference's DataLoader loop (src/train.py:158-192:
    ``coords.to(device); gt.to(device); ... backward(); optim.step()``) for callers that keep their own loader.

    A staging set is one pinned host block [coords | targets | row mask] and its device twin.  Per step: the block goes
    host -> HBM as ONE copy on a copy stream (it overlaps the previous step's kernels), then one CUDA-graph launch on
    the caller's stream runs the fused step's kernels and the D2H copy of the step's loss into pinned memory.
    ``depth`` staging sets rotate, so the host fills and launches step i+1 while step i runs; ``loss(ticket)`` waits
    for exactly that step and stays valid until the set is reused (``depth`` submits later).

    ``step_fn(coords_dev, gt_dev, mask_dev, slot, bs)`` launches the step's kernels on the current stream and leaves
    the loss in ``eng.loss_out`` (default: ``eng.train_step``)."""

    def __init__(self, eng, loss: str, batch_size: int, masked: bool = False, loss_opts: Optional[dict] = None,
                 depth: int = 2, step_fn=None, out: Optional[torch.Tensor] = None, use_graph: bool = True):
        if loss not in FUSABLE_LOSSES:
            raise L.InrError(f"loss '{loss}' is not fused; use the unfused model(x)/backward path")
        dev = eng.params.device
        self.eng, self.loss_name, self.loss_opts, self.bs, self.depth = eng, loss, loss_opts, int(batch_size), int(depth)
        in_f, out_f, bs = 3, int(eng.plan.net["network_output_size"]), self.bs     # targets: one [bs, out] block (multi-head models too)
        up = lambda n: (n + 255) // 256 * 256
        o_gt = up(bs * in_f * 4)
        o_mask = o_gt + up(bs * out_f * 4)
        total = o_mask + (up(bs) if masked else 0)

        def views(block):
            c = block[:bs * in_f * 4].view(torch.float32).view(bs, in_f)
            y = block[o_gt:o_gt + bs * out_f * 4].view(torch.float32).view(bs, out_f)
            m = block[o_mask:o_mask + bs] if masked else None
            return c, y, m
        self.h_block = [torch.zeros(total, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.d_block = [torch.zeros(total, dtype=torch.uint8, device=dev) for _ in range(depth)]
        self._h = [views(b) for b in self.h_block]
        self._d = [views(b) for b in self.d_block]
        self.h_loss = [torch.zeros(1).pin_memory() for _ in range(depth)]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._in = [torch.cuda.Event() for _ in range(depth)]     # staging set landed in HBM
        self._done = [None] * depth                                # event recorded after the set's last step
        self._graphs = {}
        self._warm = set()
        self._out = out
        self.use_graph = use_graph
        self.count = 0
        self.h2d_bytes = total
        self._step_fn = step_fn or self._default_step

    def _default_step(self, c, y, m, slot, bs):
        self.eng.train_step(self.loss_name, c, y, bs, mask=m, loss_opts=self.loss_opts, use_cursor=False, out=self._out)

    def staging(self, slot: int):
        """Pinned (coords, targets, mask) views of a staging set; valid to write once ``wait_slot(slot)`` returned."""
        return self._h[slot]

    def next_slot(self) -> int:
        return self.count % self.depth

    def wait_slot(self, slot: int):
        ev = self._done[slot]
        if ev is not None:
            ev.synchronize()

    def _body(self, slot, bs):
        c, y, m = self._d[slot]
        self._step_fn(c, y, m, slot, bs)
        self.h_loss[slot].copy_(self.eng.loss_out.reshape(-1)[:1], non_blocking=True)

    def launch(self, slot: Optional[int] = None, bs: Optional[int] = None) -> int:
        """Run one step on the batch sitting in staging set ``slot`` (the caller has waited for the set and filled it);
        returns the ticket (= slot)."""
        slot = self.next_slot() if slot is None else slot
        bs = self.bs if bs is None else int(bs)
        main = torch.cuda.current_stream()
        # the set's previous step has finished (wait_slot), so its device twin is free: no device-side wait needed
        with torch.cuda.stream(self._copy_stream):
            self.d_block[slot].copy_(self.h_block[slot], non_blocking=True)
            self._in[slot].record(self._copy_stream)
        main.wait_event(self._in[slot])
        key = (slot, bs)
        g = self._graphs.get(key) if self.use_graph else None
        if g is not None:
            g.replay()
        else:
            self._body(slot, bs)                       # eager: first uses of this (set, batch size), or graphs off
            if self.use_graph and key in self._warm:
                # second eager use done: calibration passes and kernel attributes are behind us -> record the graph
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):              # capture only records; nothing executes here
                    self._body(slot, bs)
                self._graphs[key] = g
            self._warm.add(key)
        ev = self._done[slot] or torch.cuda.Event()
        ev.record(main)
        self._done[slot] = ev
        self.count += 1
        return slot

    def submit(self, coords: torch.Tensor, gt: torch.Tensor, mask: Optional[torch.Tensor] = None) -> int:
        """Copy a host batch (CPU tensors, short last batch allowed) into the next staging set and launch its step."""
        slot = self.next_slot()
        bs = int(coords.shape[0])
        if bs > self.bs:
            raise L.InrError(f"batch of {bs} rows exceeds the stepper's batch size {self.bs}")
        hc, hy, hm = self._h[slot]
        if hm is not None and mask is None:
            raise L.InrError("this stepper was built with masked=True: every batch needs its row mask")
        self.wait_slot(slot)
        hc[:bs].copy_(coords)
        hy[:bs].copy_(gt)
        if hm is not None:
            hm[:bs].copy_(mask.reshape(-1)[:bs] if mask.dim() == 1 else mask[:, 0])
        return self.launch(slot, bs)

    def loss(self, ticket: int) -> float:
        """Loss of the step launched with this ticket (blocks until that step's D2H copy has landed)."""
        self.wait_slot(ticket)
        return float(self.h_loss[ticket])


class RingTrainer:
    """One ring model of the partitioned k-space fit (reference src/train_variations/train_clustering.py:171-189): the
    model sees every grid-order batch but only the rows whose distance to the k-space centre lies in [r0, r1] enter the
    loss -- exactly the reference's ``coords[ind]`` / ``gt[ind]`` gather, expressed as the fused step's row mask (masked-out
    rows get zero gradient, the loss is the mean over the ring's rows).  Ring models are independent of each other:
    one RingTrainer per ring, rings spread over GPUs by ``parallel.owned_rings`` with no collective while training."""

    def __init__(self, model: FusedChain, encoder: Positional_Encoder, optim: FusedAdam, loss: str, batch_size: int,
                 coords: torch.Tensor, gt: torch.Tensor, dist: torch.Tensor, loss_opts: Optional[dict] = None):
        if loss not in ("L2", "L1", "MSLE", "tanh"):
            # HDR: the reference hands the ENCODED batch to the loss as k-space coordinates (:184); LSL means CenterLoss
            # there (:85-86, randperm based) -- neither is a fused-kernel target
            raise NotImplementedError(f"ring models are fitted with L2 / L1 / MSLE / tanh on the fused path, not '{loss}'")
        if encoder.embedding_type != "gauss":
            raise L.InrError("the fused step needs the gauss encoder")
        dev = model._flat.device
        self.model, self.encoder, self.optim = model, encoder, optim
        self.loss, self.loss_opts, self.bs = loss, dict(loss_opts or {}), int(batch_size)
        self.coords = coords.to(dev, torch.float32).contiguous()
        self.gt = gt.to(dev, torch.float32).contiguous()
        self.dist = dist.to(dev, torch.float32).contiguous()
        self._dist_host = dist.detach().to("cpu", torch.float32).numpy()
        self.n = self.coords.shape[0]
        self.eng = model.engine(encoder.params, self.bs)
        self.eng.set_encoder(encoder.B)

    @property
    def steps_per_epoch(self):
        return (self.n + self.bs - 1) // self.bs

    def step(self, it: int, r0: float, r1: float):
        """Batch `it` (grid order) restricted to the ring [r0, r1].  Returns the device scalar holding the loss, or None
        when no row of the batch lies in the ring (reference :179: the model is not touched)."""
        import numpy as np
        i = it * self.bs
        bs = min(self.bs, self.n - i)
        dh = self._dist_host[i:i + bs]
        if not bool(((dh >= np.float32(r0)) & (dh <= np.float32(r1))).any()):      # host copy: no device sync
            return None
        d = self.dist[i:i + bs]
        mask = ((d >= r0) & (d <= r1)).to(torch.uint8)
        self.optim.sync_hyper(self.eng)
        self.model.engine(self.encoder.params, self.bs)
        self.eng.train_step(self.loss, self.coords[i:i + bs], self.gt[i:i + bs], bs, mask=mask, loss_opts=self.loss_opts,
                            use_cursor=False)
        self.model.mark_params_updated_by_kernel(self.eng)
        self.optim._fused_done = False
        return self.eng.loss_out

    @torch.no_grad()
    def predict_rows(self, rows: torch.Tensor, coords: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused forward on the given row indices of `coords` (default: the resident training coordinates) -- the
        validation of one ring (reference :218-232)."""
        coords = self.coords if coords is None else coords
        self.model.engine(self.encoder.params, self.bs)      # refreshes the fp16 operand copies after load_state_dict
        out = torch.empty(rows.numel(), self.eng.plan.out_cols, dtype=torch.float32, device=self.coords.device)
        for s in range(0, rows.numel(), self.bs):
            c = coords[rows[s:s + self.bs]].contiguous()
            out[s:s + c.shape[0]] = self.eng.forward(c, train=False)
        return out
