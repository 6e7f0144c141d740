"""Host-side engine: flat fp32 parameter / Adam buffers, fp16 operand copies, workspace, CUDA graphs.

Mirrors the per-batch body of the reference training loop (src/train.py:158-192): every public method
is a thin wrapper that turns torch tensors into raw device pointers for one C-ABI call."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _loss_desc(loss: str, loss_opts: Optional[dict]):
    """inr_loss_desc from the reference's loss name + loss_opts; loss_opts['tv'] = (H, W[, weight]) switches the
    total-variation term of the per-coil loop on (src/train.py:173-174; weight defaults to tv_loss's 1e-4)."""
    o = loss_opts or {}
    tv = o.get("tv")
    tvw, tvh, tvww = (float(tv[2]) if len(tv) > 2 else 1e-4, int(tv[0]), int(tv[1])) if tv else (0.0, 0, 0)
    dp = o.get("dp_norm")            # (device float table [n_batches, 2], rows per local batch): see inr_loss_desc.dp_norm
    dp_ptr, dp_rows = (dp[0].data_ptr(), int(dp[1])) if dp is not None else (None, 0)
    d = L.LossDesc(L.LOSS[loss], float(o.get("hdr_eps", 0.0)), float(o.get("hdr_ff_sigma", 1.0)),
                   float(o.get("hdr_ff_factor", 0.0)), tvw, tvh, tvww, dp_ptr, dp_rows)
    cons = o.get("consistency")      # (bounds [(lo, hi), ...], weight): ConsistencyLoss of the multi-scale loop
    if cons is not None:
        bounds, weight = cons
        if len(bounds) > 8:
            raise L.InrError("at most 8 consistency bounds")
        d.cons_weight = float(weight)
        for i, (lo, hi) in enumerate(bounds):
            d.cons_bounds[2 * i], d.cons_bounds[2 * i + 1] = float(lo), float(hi)
    return d


def selftest_umma(mode: int, variant: int = 0):
    err, ref = C.c_float(), C.c_float()
    L.check(L.lib.inr_selftest_umma(mode, variant, C.byref(err), C.byref(ref)), "inr_selftest_umma")
    return err.value, ref.value


def set_sm_budget(n_sm: int):
    """Cap every chained (WIRE / WIRE2D layer-chain) launch of this process at `n_sm` SMs so that several fits can step
    concurrently on one device without starving each other (inr_set_sm_budget); 0 = the whole chip."""
    L.check(L.lib.inr_set_sm_budget(int(n_sm)), "inr_set_sm_budget")


class Plan:
    """Immutable description of one model (inr_plan): layer table, parameter layout, work units."""

    def __init__(self, model: str, net: dict, encoder: Optional[dict] = None):
        enc = encoder or {"embedding": "none"}
        if model not in L.MODEL:
            raise NotImplementedError(model)                      # src/train.py:69-70
        if model == "WIRE2D" and net.get("last_tanh", False):
            last = "tanh"                                          # src/models/wire2d.py:106-107: complex tanh, then .real
        elif model in ("WIRE", "WIRE2D", "Fourier", "MultiscaleFourier", "BoundedFourier", "Gabor", "KGabor"):
            last = "linear"                                        # WIRE: real part of the final complex linear; MFN: plain heads
        elif model == "FFN":
            last = "sigmoid"                                       # src/models/networks.py:63
        elif net.get("last_tanh", False):
            last = "tanh"                                          # src/models/networks.py:94-95
        elif not net.get("network_last_linear", True):
            last = "sin"                                           # src/models/networks.py:107-117: sine output layer
        else:
            last = "linear"
        kind = enc.get("embedding", "none")
        if kind not in L.ENC:
            raise L.InrError(f"encoder '{kind}' is not built into the fused kernels")
        if kind == "LogF" and model not in ("SIREN", "FFN"):
            raise L.InrError("the LogF encoder is fused for SIREN / FFN only (src/models/networks.py:24-29)")
        self.desc = L.ModelDesc(L.MODEL[model], int(net["network_input_size"]), int(net["network_output_size"]),
                                int(net["network_depth"]), int(net["network_width"]), L.LAST[last],
                                L.ENC[kind], int(enc.get("embedding_size", 0)) if kind in ("gauss", "LogF") else 0,
                                float(net.get("first_omega_0", 30.0)) if model in ("WIRE", "WIRE2D") else 30.0,
                                float(net.get("hidden_omega_0", 30.0)), float(net.get("scale", 10.0)))
        self.out_cols = int(net["network_output_size"])
        # SIREN / FFN outside the on-chip chain kernels' shape (width 256, >= 2 sine layers, linear / tanh / sigmoid output)
        # run layer by layer on the streaming stage GEMMs ("wide chain", csrc/abi.cu wide_plan_create)
        # (and so does every LogF-encoded model: its 6 n input features are padded to the operand chunks there)
        self.wide = model in ("SIREN", "FFN") and (int(net["network_width"]) != 256 or int(net["network_depth"]) < 3 or last == "sin"
                                                    or kind == "LogF")
        if model in ("MultiscaleFourier", "BoundedFourier"):
            layers = list(net.get("output_layers", [1, 3, 5, 7]))
            mask = 0
            for i in layers:
                mask |= 1 << int(i)
            self.desc.head_mask = mask
            self.out_cols = len(layers) * int(net["network_output_size"])
            for i, (lo, hi) in enumerate(net.get("boundaries", []) or []):
                self.desc.bounds[2 * i], self.desc.bounds[2 * i + 1] = float(lo), float(hi)
        self.model, self.net, self.encoder = model, dict(net), dict(enc)
        h = C.c_void_p()
        L.check(L.lib.inr_plan_create(C.byref(self.desc), C.byref(h)), "inr_plan_create")
        self.handle = h
        n = C.c_int64()
        L.check(L.lib.inr_plan_param_count(h, C.byref(n)), "inr_plan_param_count")
        self.n_params = n.value
        nt = C.c_int32()
        L.check(L.lib.inr_plan_tensor_count(h, C.byref(nt)), "inr_plan_tensor_count")
        self.tensors, self.tensor_flags = [], []
        for i in range(nt.value):
            ti = L.TensorInfo()
            L.check(L.lib.inr_plan_tensor(h, i, C.byref(ti)), "inr_plan_tensor")
            self.tensors.append((ti.offset, ti.rows, ti.cols, ti.layer, bool(ti.is_bias)))
            self.tensor_flags.append((bool(ti.is_complex), bool(ti.frozen)))
        b = C.c_size_t()
        L.check(L.lib.inr_wpack_bytes(h, C.byref(b)), "inr_wpack_bytes")
        self.wpack_bytes = b.value

    def workspace_bytes(self, bs: int) -> int:
        b = C.c_size_t()
        L.check(L.lib.inr_workspace_bytes(self.handle, bs, C.byref(b)), "inr_workspace_bytes")
        return b.value

    def scalars_offset(self, bs: int) -> int:
        b = C.c_size_t()
        L.check(L.lib.inr_scalars_offset(self.handle, bs, C.byref(b)), "inr_scalars_offset")
        return b.value

    def workspace_layout(self, bs: int) -> dict:
        arr = (C.c_uint64 * 60)()
        L.check(L.lib.inr_workspace_layout(self.handle, bs, arr, 60), "inr_workspace_layout")
        v = list(arr)
        return {"h": v[0:12], "d": v[12:24], "dz": v[24:36], "dzlast": v[36], "g": v[37], "part": v[38],
                "scal": v[39], "gpart": v[40], "n_tiles": int(v[41]), "n_split": int(v[42]), "total": v[43],
                "q": v[44:56], "gfin": v[56], "gstride": int(v[57]), "aux0": int(v[58]), "e": v[59],
                # width-256 chains: rows per tile and k-group stride of the operand images (chain_t.cu below one wave of rows)
                "tile_rows": int(v[44]) if self.model in ("SIREN", "FFN") and not self.wide else 128,
                "lb": int(v[45]) if self.model in ("SIREN", "FFN") and not self.wide else 2048}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and getattr(L, "lib", None) is not None:
            L.lib.inr_plan_destroy(h)
            self.handle = None


class ChainEngine:
    """Owns the device state of one fit: parameters, Adam moments, fp16 operand copies, workspace."""

    def __init__(self, plan: Plan, max_batch: int, device="cuda", lr=5e-4, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, reg_l1=0.0, reg_l2=0.0, shared: Optional[dict] = None):
        if not torch.cuda.is_available():
            raise L.InrError("ChainEngine needs a CUDA device (no CPU fallback)")
        self.plan, self.device, self.max_batch = plan, torch.device(device), int(max_batch)
        P = plan.n_params
        f32 = dict(dtype=torch.float32, device=self.device)
        sh = shared or {}                         # engines of one module share parameters / Adam state
        self.params = sh.get("params", None) if sh else None
        if self.params is None:
            self.params = torch.zeros(P, **f32)
        assert self.params.numel() == P and self.params.is_cuda and self.params.dtype == torch.float32
        self.grads = sh["grads"] if "grads" in sh else torch.zeros(P, **f32)
        self.exp_avg = sh["exp_avg"] if "exp_avg" in sh else torch.zeros(P, **f32)
        self.exp_avg_sq = sh["exp_avg_sq"] if "exp_avg_sq" in sh else torch.zeros(P, **f32)
        self.wpack = torch.zeros(plan.wpack_bytes + 1024, dtype=torch.uint8, device=self.device)
        self.workspace = torch.zeros(plan.workspace_bytes(self.max_batch), dtype=torch.uint8, device=self.device)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, reg_l1, reg_l2, 0.0], **f32)
        self.step = sh["step"] if "step" in sh else torch.zeros(1, dtype=torch.int32, device=self.device)
        self.cursor = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.loss_out = torch.zeros(1, **f32)
        self.encB = None
        self._graphs = {}
        # WIRE / MFN backward passes store fp16 gradient images under per-layer power-of-two scales that lag one step
        # (amax of the previous backward); the very first backward after construction runs twice so that the gradients
        # it returns are already computed with calibrated scales
        self._lagged_scales = plan.model not in ("SIREN", "FFN") or plan.wide
        self._calibrated = False

    # ---- parameters -------------------------------------------------------------------------------
    def _views(self, flat):
        out = []
        for (off, rows, cols, layer, is_bias), (is_complex, frozen) in zip(self.plan.tensors, self.plan.tensor_flags):
            if is_complex:      # interleaved (re, im) pairs == torch complex64 storage
                v = torch.view_as_complex(flat[off:off + rows * cols * 2].view(rows * cols, 2))
            else:
                v = flat[off:off + rows * cols]
            out.append(v if (is_bias or (rows == 1 and cols == 1)) else v.view(rows, cols))
        return out

    def param_views(self):
        """Views into the flat buffer, one per reference tensor, in state_dict order."""
        return self._views(self.params)

    def grad_views(self):
        return self._views(self.grads)

    def load_tensors(self, tensors):
        """Copy a list of tensors (reference state_dict values, same order) into the flat buffer and repack."""
        views = self.param_views()
        assert len(tensors) == len(views), (len(tensors), len(views))
        with torch.no_grad():
            for v, t in zip(views, tensors):
                v.copy_(t.to(self.device, v.dtype).reshape(v.shape))
        self.pack()

    def set_encoder(self, B: Optional[torch.Tensor]):
        self.encB = None if B is None else B.to(self.device, torch.float32).contiguous()

    def set_lr(self, lr: float):
        self.hyper[0:1].fill_(lr)

    def pack(self):
        L.check(L.lib.inr_pack_weights(self.plan.handle, _ptr(self.params), _ptr(self.wpack), _stream()), "inr_pack_weights")

    # ---- unfused API (keeps arbitrary PyTorch losses working) --------------------------------------
    def forward(self, inp: torch.Tensor, train: bool = False, dist: Optional[torch.Tensor] = None) -> torch.Tensor:
        bs = inp.shape[0]
        assert bs <= self.max_batch
        inp = inp.to(self.device, torch.float32).contiguous()
        out = torch.zeros(bs, self.plan.out_cols, dtype=torch.float32, device=self.device)
        if dist is not None:
            dist = dist.to(self.device, torch.float32).contiguous()
            L.check(L.lib.inr_forward_dist(self.plan.handle, _ptr(self.params), _ptr(self.wpack), _ptr(inp), _ptr(self.encB),
                                           _ptr(dist), bs, _ptr(self.workspace), _ptr(out), 1 if train else 0, _stream()),
                    "inr_forward_dist")
            return out
        L.check(L.lib.inr_forward(self.plan.handle, _ptr(self.params), _ptr(self.wpack), _ptr(inp), _ptr(self.encB), bs,
                                  _ptr(self.workspace), _ptr(out), 1 if train else 0, _stream()), "inr_forward")
        return out

    def backward(self, dout: torch.Tensor, dist: Optional[torch.Tensor] = None) -> torch.Tensor:
        bs = dout.shape[0]
        dout = dout.to(self.device, torch.float32).contiguous()
        if dist is not None:
            dist = dist.to(self.device, torch.float32).contiguous()
            for _ in range(2 if (self._lagged_scales and not self._calibrated) else 1):
                L.check(L.lib.inr_backward_dist(self.plan.handle, _ptr(self.params), _ptr(self.wpack), _ptr(dout), _ptr(dist), bs,
                                                _ptr(self.workspace), _ptr(self.grads), _stream()), "inr_backward_dist")
            self._calibrated = True
            return self.grads
        for _ in range(2 if (self._lagged_scales and not self._calibrated) else 1):
            L.check(L.lib.inr_backward(self.plan.handle, _ptr(self.params), _ptr(self.wpack), _ptr(dout), bs,
                                       _ptr(self.workspace), _ptr(self.grads), _stream()), "inr_backward")
        self._calibrated = True
        return self.grads

    def adam_step(self):
        self.step += 1
        L.check(L.lib.inr_adam_step(self.plan.handle, _ptr(self.params), _ptr(self.grads), _ptr(self.exp_avg),
                                    _ptr(self.exp_avg_sq), _ptr(self.wpack), _ptr(self.hyper), _ptr(self.step), _stream()),
                "inr_adam_step")

    def adam_step_peers(self, ex, parity: int):
        """Data-parallel optimiser step with the gradient exchange fused into the kernel: gradients = mean over ranks
        of the peer-mapped buffers of `ex` (a parallel.PeerGradExchange) for the given parity, which must alternate with
        every optimiser step (step & 1: the same buffer `grad_step` of this step wrote) -- see PeerGradExchange.
        No rank may do long host-only work between the steps of a fit without the others: the entry barrier of the
        exchange waits for every rank (INR_PEER_TIMEOUT_S, default 120 s, then the kernel traps).  The kernel advances the
        device step counter itself."""
        gp, fp = ex.pointers(parity)
        L.check(L.lib.inr_adam_step_peers(self.plan.handle, _ptr(self.params), gp, fp, ex.world, ex.rank, _ptr(self.exp_avg),
                                          _ptr(self.exp_avg_sq), _ptr(self.wpack), _ptr(self.hyper), _ptr(self.step), _stream()),
                "inr_adam_step_peers")

    # ---- fused step --------------------------------------------------------------------------------
    def train_step(self, loss: str, coords: Optional[torch.Tensor], gt: torch.Tensor, bs: int, x: Optional[torch.Tensor] = None,
                   mask: Optional[torch.Tensor] = None, loss_opts: Optional[dict] = None, use_cursor: bool = False,
                   out: Optional[torch.Tensor] = None, dist: Optional[torch.Tensor] = None):
        """One fused forward+loss+backward+Adam step on rows [cursor, cursor+bs) (or [0, bs)) of the resident
        arrays.  Asynchronous; the loss lands in self.loss_out.  dist (multi-scale models): fp32 [rows] distance to the
        k-space centre, indexed like coords; loss_opts['consistency'] = (bounds, weight) adds the ConsistencyLoss term."""
        ld = _loss_desc(loss, loss_opts)
        if dist is not None or self.plan.model in ("MultiscaleFourier", "BoundedFourier"):
            if not self._calibrated:        # gradients-only pass: calibrates the lagged per-stage gradient scales
                L.check(L.lib.inr_grad_step_dist(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.wpack), _ptr(coords),
                                                 _ptr(x), _ptr(self.encB), _ptr(gt), _ptr(mask), _ptr(dist), bs, None,
                                                 _ptr(self.workspace), _ptr(out), _ptr(self.grads), _ptr(self.loss_out), _stream()),
                        "inr_grad_step_dist(calibration)")
            L.check(L.lib.inr_train_step_dist(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.exp_avg),
                                              _ptr(self.exp_avg_sq), _ptr(self.wpack), _ptr(self.hyper), _ptr(self.step),
                                              _ptr(coords), _ptr(x), _ptr(self.encB), _ptr(gt), _ptr(mask), _ptr(dist), bs,
                                              _ptr(self.cursor) if use_cursor else None, _ptr(self.workspace), _ptr(out),
                                              _ptr(self.loss_out), _stream()), "inr_train_step_dist")
            self._calibrated = True
            return
        if self._lagged_scales and not self._calibrated:
            self._calibrate(ld, coords, x, gt, mask, bs, out)
        L.check(L.lib.inr_train_step(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                     _ptr(self.wpack), _ptr(self.hyper), _ptr(self.step), _ptr(coords), _ptr(x), _ptr(self.encB),
                                     _ptr(gt), _ptr(mask), bs, _ptr(self.cursor) if use_cursor else None, _ptr(self.workspace),
                                     _ptr(out), _ptr(self.loss_out), _stream()), "inr_train_step")

    def _calibrate(self, ld, coords, x, gt, mask, bs, out=None):
        """One gradients-only pass over rows [0, bs) (no optimiser step, cursor and step counter untouched) that leaves
        the per-layer gradient scales of the WIRE / MFN backward calibrated for the step that follows."""
        L.check(L.lib.inr_grad_step(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.wpack), _ptr(coords), _ptr(x),
                                    _ptr(self.encB), _ptr(gt), _ptr(mask), bs, None, _ptr(self.workspace), _ptr(out),
                                    _ptr(self.grads), _ptr(self.loss_out), _stream()), "inr_grad_step(calibration)")
        self._calibrated = True

    def profile_step(self, loss: str, coords, gt, bs: int, x=None, mask=None, loss_opts=None, reps: int = 20, out=None):
        """Average device time (ms) of the four kernels of one step: forward, dgrad, wgrad, optimiser."""
        ld = _loss_desc(loss, loss_opts)
        ms = (C.c_float * 8)()
        L.check(L.lib.inr_profile_step(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.exp_avg),
                                       _ptr(self.exp_avg_sq), _ptr(self.wpack), _ptr(self.hyper), _ptr(self.step), _ptr(coords),
                                       _ptr(x), _ptr(self.encB), _ptr(gt), _ptr(mask), bs, _ptr(self.workspace), _ptr(out),
                                       reps, ms, _stream()), "inr_profile_step")
        if self.plan.model in ("WIRE", "WIRE2D"):
            return {"forward": ms[0], "backward": ms[2], "optimiser": ms[3], "forward_layer_gemms": ms[4]}
        if self.plan.model not in ("SIREN", "FFN") or self.plan.wide:      # MFN family / wide chains: forward (+ TV) | backward (dgrad + wgrad) | optimiser
            return {"forward": ms[0], "backward": ms[2], "optimiser": ms[3]}
        return {"forward": ms[0], "dgrad": ms[1], "wgrad": ms[2], "optimiser": ms[3]}

    def read_image(self, kind: str, layer: int, bs: int) -> torch.Tensor:
        """Decode one saved fp16 operand image family back to a [rows_pad, F] fp32 matrix (tests / debugging).
        kind: 'h' (input of `layer`), 'd' (act'(z) of `layer`), 'dz' (dZ of `layer`), 'dzlast'."""
        lay = self.plan.workspace_layout(bs)
        T = lay["n_tiles"]
        n_gemm = self.plan.desc.depth - 1
        if kind == "dzlast":
            F, off = 16, lay["dzlast"]
        else:
            F = self.plan.desc.in_features if (kind == "h" and layer == 0) else self.plan.desc.width
            off = lay[kind][layer]
        R, lb = lay["tile_rows"], lay["lb"]           # elem(r, f) of a tile at (f/8)*lb + r*16 + (f%8)*2
        nbytes = T * (F // 8) * lb
        img = self.workspace[off:off + nbytes].view(torch.float16).view(T, F // 8, lb // 16, 8)[:, :, :R]
        mat = img.permute(0, 2, 1, 3).reshape(T * R, F).float()
        if kind == "h" and layer == 0 and self.plan.desc.encoder == L.ENC["gauss"]:
            E = self.plan.desc.enc_size           # undo the sin/cos chunk interleave of the in-kernel encoder
            kp = torch.arange(F, device=mat.device)
            c64, kk = kp // 64, kp % 64
            real = torch.where(kk < 32, 32 * c64 + kk, E + 32 * c64 + kk - 32)
            out = torch.empty_like(mat)
            out[:, real] = mat
            mat = out
        return mat

    def grad_step(self, loss: str, coords, gt, bs: int, x=None, mask=None, loss_opts=None, use_cursor: bool = False, out=None,
                  grads: Optional[torch.Tensor] = None, dist: Optional[torch.Tensor] = None):
        """forward + loss + backward only; gradients land in self.grads, or in `grads` (e.g. a PeerGradExchange buffer)
        (data-parallel: all-reduce them, then adam_step -- or adam_step_peers)."""
        gbuf = self.grads if grads is None else grads
        assert gbuf.numel() >= self.plan.n_params and gbuf.dtype == torch.float32 and gbuf.is_cuda
        ld = _loss_desc(loss, loss_opts)
        if dist is not None or self.plan.model in ("MultiscaleFourier", "BoundedFourier"):
            for _ in range(1 if self._calibrated else 2):
                L.check(L.lib.inr_grad_step_dist(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.wpack), _ptr(coords),
                                                 _ptr(x), _ptr(self.encB), _ptr(gt), _ptr(mask), _ptr(dist), bs,
                                                 _ptr(self.cursor) if use_cursor else None, _ptr(self.workspace), _ptr(out),
                                                 _ptr(gbuf), _ptr(self.loss_out), _stream()), "inr_grad_step_dist")
            self._calibrated = True
            return gbuf
        if self._lagged_scales and not self._calibrated:
            self._calibrate(ld, coords, x, gt, mask, bs, out)
        L.check(L.lib.inr_grad_step(self.plan.handle, C.byref(ld), _ptr(self.params), _ptr(self.wpack), _ptr(coords), _ptr(x),
                                    _ptr(self.encB), _ptr(gt), _ptr(mask), bs, _ptr(self.cursor) if use_cursor else None,
                                    _ptr(self.workspace), _ptr(out), _ptr(gbuf), _ptr(self.loss_out), _stream()),
                "inr_grad_step")
        return gbuf

    def read_mfn_image(self, kind: str, stage: int, bs: int) -> torch.Tensor:
        """MFN only (tests / debugging): 'z' (stage output), 'g' (sin p), 'dp' (S_stage * dL/dp) as [rows_pad, width]."""
        lay = self.plan.workspace_layout(bs)
        T, F = lay["n_tiles"], self.plan.desc.width
        if kind == "x":         # the encoded input image [rows_pad, in_features padded to 128s] (slot 36 of the MFN layout)
            F = (self.plan.desc.in_features + 127) // 128 * 128
            off = lay["dzlast"]
        else:
            off = {"z": lay["h"], "g": lay["d"], "dp": lay["dz"], "q": lay["q"]}[kind][stage]
        img = self.workspace[off:off + T * 128 * F * 2].view(torch.float16).view(T, F // 8, 128, 8)
        return img.permute(0, 2, 1, 3).reshape(T * 128, F).float()

    def read_wire_image(self, kind: str, layer: int, bs: int) -> torch.Tensor:
        """WIRE only (tests / debugging): decode a 384-feature image family to a complex [rows_pad, 192] matrix.
        kind: 'h' (hi + lo parts of the complex input of `layer`), 'ab' (pre-activation a + jb of `layer`),
        'dz' (S * dL/d(a + jb) of `layer`)."""
        lay = self.plan.workspace_layout(bs)
        T = lay["n_tiles"]
        nbytes = T * 128 * 384 * 2

        def dec(off):
            img = self.workspace[off:off + nbytes].view(torch.float16).view(T, 48, 128, 8)
            return img.permute(0, 2, 1, 3).reshape(T * 128, 384).float()

        if kind == "h":
            hi = dec(lay["h"][layer])
            lo = dec(lay["h"][layer] + nbytes)          # H_lo[l] is laid out right after H_hi[l]
            # WIRE: the lo image of the last hidden layer's output is never written (nothing reads it)
            last = self.plan.model == "WIRE" and layer == int(self.plan.net["network_depth"]) + 1
            m = hi if last else hi + lo
        elif kind == "ab":
            m = dec(lay["d"][layer])
        else:
            m = dec(lay["dz"][layer])
        return torch.complex(m[:, :192], m[:, 192:])

    def scalars(self, bs: int) -> torch.Tensor:
        off = self.plan.scalars_offset(bs)
        return self.workspace[off:off + 256].view(torch.float32).clone()
