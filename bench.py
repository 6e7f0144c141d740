#!/usr/bin/env python
"""Benchmark of the INR fitting hot path (forward + loss + backward + Adam) -- training coordinates / second.

    python bench.py --gpus N --steps K --warmup W              # this engine, N GPUs of one node
    python bench.py --impl reference --steps K --warmup W      # the reference's CPU path (oracle port), host cores

One "step" = one grid-order batch of the configured size through the whole hot path.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
    # BASELINE.json configs[1] (the configuration the metric is quoted on): WIRE complex Gabor depth 4 width 256
    # (-> 181 complex), k-space fit, HDR loss, batch 25000, undersampling grid-2*1 (every other row enters the loss)
    "wire_kspace_hdr_bs25000": dict(
        model="WIRE", loss="HDR", loss_opts={"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}, batch=25000,
        image_space=False, normalization="max", undersampling="grid-2*1",
        net={"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256,
             "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15},
        encoder={"embedding": "none", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=3155916, fwd_flop_per_coord=4 * 181 * 181 * 8, issued_fwd_flop_per_coord=4 * 3 * 2 * 384 * 384),
    # BASELINE.json configs[0]: SIREN depth 4 width 256, gauss-512 encoding, image space, L2, batch 10000
    "siren_image_l2_bs10000": dict(
        model="SIREN", loss="L2", loss_opts=None, batch=10000, image_space=True, normalization="max", undersampling=None,
        net={"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256},
        encoder={"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=1313792, fwd_flop_per_coord=525312),
    # batch sweep of configs[0] (the reference's config_siren_image.yaml fits with batch_size 300000)
    **{f"siren_image_l2_bs{b}": dict(
        model="SIREN", loss="L2", loss_opts=None, batch=b, image_space=True, normalization="max", undersampling=None,
        net={"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256},
        encoder={"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=1313792, fwd_flop_per_coord=525312) for b in (25000, 100000, 300000)},
    # the reference's SIREN k-space configs (config_siren_kspace*.yaml): width 512, depth 8, last_tanh, L2, batch 100000
    "siren_w512_d8_kspace_l2_bs100000": dict(
        model="SIREN", loss="L2", loss_opts=None, batch=100000, image_space=False, normalization="max", undersampling=None,
        net={"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 512, "last_tanh": True},
        encoder={"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=3 * 2 * (512 * 512 + 6 * 512 * 512 + 512 * 2) - 2 * 512 * 512, fwd_flop_per_coord=2 * (512 * 512 + 6 * 512 * 512 + 512 * 2)),
    # WIRE2D (reference config_wire2d_kspace.yaml: depth 8, width 256, omega 30, scale 15), k-space fit, L2
    "wire2d_kspace_l2_bs25000": dict(
        model="WIRE2D", loss="L2", loss_opts=None, batch=25000, image_space=False, normalization="max", undersampling=None,
        net={"network_input_size": 3, "network_output_size": 2, "network_depth": 8, "network_width": 256,
             "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15},
        encoder={"embedding": "none", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=6144 + 8 * 3145728 + 12288, fwd_flop_per_coord=8 * 2 * 256 * 256 * 8,
        issued_fwd_flop_per_coord=8 * 3 * 2 * 512 * 1024),
    # BASELINE.json configs[2] (Gabor arm): GaborNet depth 8 width 512 on gauss-512, k-space fit, tanh loss, per-coil
    # batches (bs = H*W = 102400), undersampling grid-2*1, total-variation term on every batch
    "gabor_kspace_tanh_tv_percoil": dict(
        model="Gabor", loss="tanh", loss_opts={"tv": (320, 320)}, batch=102400, image_space=False, normalization="max",
        undersampling="grid-2*1",
        net={"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 512},
        encoder={"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
        flop_per_coord=31463424, fwd_flop_per_coord=(18 + 8) * 512 * 512 * 2 + 2 * 512 * 2),
}
# BASELINE.json configs[3]: MultiscaleBoundedFourier depth 8 width 512 (4 ring heads), LSL loss per head + 0.1 *
# ConsistencyLoss, batch 100000, non-per-coil: the loop body of src/train_kspace_multiscale.py as one fused CUDA-graph step
# (inr_train_step_dist).  Measured by run_multiscale() below (single GPU).
MULTISCALE_WORKLOAD = dict(
    name="bounded_fourier_lsl_bs100000", model="BoundedFourier", batch=100000, image_space=False, normalization="max",
    net={"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 512},
    encoder={"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3},
    loss_opts={"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.0}, radii=[0.35, 0.7, 1.05, 5.0],
    flop_per_coord=19423232)
DEFAULT_WORKLOAD = "wire_kspace_hdr_bs25000"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel, read from the committed artefact of
# the latest `ncu --set full` capture (profiles/roofline_traffic.json: {workload: {bytes, source, commit}}); absent -> null


def load_traffic(workload):
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f).get(workload)
        return (float(t["bytes"]), f"{t['source']} (kernel as of commit {t['commit']})") if t else (None, None)
    except Exception:
        return None, None
SLICE = (15, 320, 320)                 # fastMRI-knee-shaped: 15 coils x 320 x 320 after the reference's crop
LR = 5e-4


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop, self.max_mhz = [], set(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def sample(self):
        """One reading; also called by the timing thread right after the last step has been enqueued, so that even a
        timed region of a few milliseconds holds a sample taken while the GPU is inside it."""
        nv = self.nv
        if not nv:
            return
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def build_engine(wl, device, seed):
    import mri_implicit_neural_representations_b200 as inr
    from mri_implicit_neural_representations_b200 import init as pinit
    torch.manual_seed(seed)
    encB = pinit.encoder_matrix(wl["encoder"])
    if wl["model"] == "WIRE":
        tensors = [t for _, t in pinit.wire_tensors(wl["net"])]
    elif wl["model"] == "WIRE2D":
        tensors = [t for _, t in pinit.wire2d_tensors(wl["net"])]
    elif wl["model"] in ("Gabor", "KGabor", "Fourier"):
        tensors = [t for _, t in pinit.mfn_tensors(wl["model"], wl["net"])]
    else:
        tensors = [t for _, t in pinit.chain_tensors(wl["model"], wl["net"])]
    plan = inr.Plan(wl["model"], wl["net"], wl["encoder"])
    eng = inr.ChainEngine(plan, max_batch=wl["batch"], device=device, lr=LR)
    eng.load_tensors(tensors)
    eng.set_encoder(encB)
    return eng, tensors, encB


def resident_arrays(wl, device, seed, min_bytes):
    """Coordinates + targets of enough synthetic slices to exceed L2 (grid-order walk = cold inputs)."""
    from mri_implicit_neural_representations_b200 import synthetic
    C, H, W = SLICE
    coords, gt = [], []
    n = 0
    while n * 20 < min_bytes:
        c, g, _ = synthetic.make_fit_arrays(seed + len(coords), C, H, W, image_space=wl["image_space"],
                                            normalization=wl["normalization"])
        coords.append(c)
        gt.append(g)
        n += c.shape[0]
    coords, gt = torch.cat(coords), torch.cat(gt)
    mask = None
    if wl["undersampling"]:          # grid-x*y: rows ::x, columns ::y are sampled (reference undersampler.py:79-91)
        gx, gy = (int(v) for v in wl["undersampling"].split("-")[1].split("*"))
        m = torch.zeros(H, W, dtype=torch.bool)
        m[::gx, ::gy] = True
        mask = m[None].expand(C, H, W).reshape(-1).repeat(coords.shape[0] // (C * H * W))
        gt = gt * mask[:, None]      # unsampled k-space points are zero-filled (undersampler.py:59-61)
    return coords.to(device), gt.to(device), (None if mask is None else mask.to(torch.uint8).to(device))


def cpu_port_steps(wl, n_steps, warmup, rows_per_step, seed=1234):
    """The reference's CPU path (oracle port: plain torch fp32 + autograd + Adam) on the host cores."""
    from oracle import inr_oracle as O
    from mri_implicit_neural_representations_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    encB = O.encoder_init(wl["encoder"])
    sd = O.MODEL_INIT[wl["model"]](dict(wl["net"]))
    C, H, W = SLICE
    coords, gt, _ = synthetic.make_fit_arrays(seed, C, H, W, image_space=wl["image_space"], normalization=wl["normalization"])
    mask = None
    if wl["undersampling"]:
        gx, gy = (int(v) for v in wl["undersampling"].split("-")[1].split("*"))
        m = torch.zeros(H, W, dtype=torch.bool)
        m[::gx, ::gy] = True
        mask = m[None].expand(C, H, W).reshape(-1)
        gt = gt * mask[:, None]
    need = (n_steps + warmup) * rows_per_step
    reps = (need + coords.shape[0] - 1) // coords.shape[0]
    if reps > 1:
        coords, gt = coords.repeat(reps, 1), gt.repeat(reps, 1)
        mask = None if mask is None else mask.repeat(reps)
    opts = None
    if wl["loss_opts"] and "hdr_eps" in wl["loss_opts"]:
        opts = {"sigma": wl["loss_opts"]["hdr_ff_sigma"], "eps": wl["loss_opts"]["hdr_eps"],
                "factor": wl["loss_opts"]["hdr_ff_factor"]}
    tv = (wl["loss_opts"] or {}).get("tv")          # only applies when a step is one whole coil (rows_per_step == H*W)
    if warmup:
        O.train_steps(wl["model"], wl["net"], sd, encB, wl["encoder"]["embedding"], coords, gt, warmup, rows_per_step, LR,
                      wl["loss"], opts, mask=mask, tv=tv)
    o = warmup * rows_per_step
    t0 = time.perf_counter()
    O.train_steps(wl["model"], wl["net"], sd, encB, wl["encoder"]["embedding"], coords[o:], gt[o:], n_steps, rows_per_step, LR,
                  wl["loss"], opts, mask=None if mask is None else mask[o:], tv=tv)
    dt = time.perf_counter() - t0
    return n_steps * rows_per_step / dt, dt


def run_reference(args, wl, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = wl["batch"]
    # bound the run: probe one step, shrink the per-step sample if K steps would take > ~150 s
    probe = min(bs, 2048)
    v1, dt1 = cpu_port_steps(wl, 1, 1, probe)
    est = dt1 * bs / probe                       # HDR / complex GEMMs scale ~linearly in the batch
    rows = bs
    if est * (args.steps + args.warmup) > 150.0:
        rows = max(256, int(bs * 150.0 / (est * (args.steps + args.warmup))) // 128 * 128)
    value, dt = cpu_port_steps(wl, args.steps, args.warmup, rows)
    cores = os.cpu_count() or 1
    sample = f"{args.steps} steps x {rows} coords of the {bs}-coord batch workload, torch CPU fp32, {torch.get_num_threads()} threads"
    line = {"impl": "reference", "metric": "train coords/sec (fwd+bwd+Adam)", "value": value, "unit": "coords/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "batch": bs, "slice": list(SLICE)},
            "cpu_baseline": {"value": value, "unit": "coords/s", "cores": cores, "kind": "port",
                             "port": "oracle restatement of the reference modules (separable HDR: faster than the reference's [bs, m] outer product)",
                             "sample": sample},
            "e2e": {"value": value, "unit": "coords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_multiscale(args):
    """BASELINE config 4 (see MULTISCALE_WORKLOAD): the loop body of src/train_kspace_multiscale.py:164-192 as ONE fused
    CUDA-graph step (FusedTrainer -> inr_train_step_dist): model(coords, dist) -> 4 x LogSpaceLoss + 0.1 ConsistencyLoss ->
    backward -> Adam."""
    wl = MULTISCALE_WORKLOAD
    sys.path.insert(0, os.path.join(ROOT, "src"))
    from mri_implicit_neural_representations_b200.modules import MultiscaleBoundedFourier, Positional_Encoder
    from mri_implicit_neural_representations_b200.trainer import FusedAdam, FusedTrainer, HostFedStepper
    from mri_implicit_neural_representations_b200 import synthetic
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    bs = wl["batch"]
    peaks = load_peaks()
    torch.manual_seed(1234)
    enc = Positional_Encoder(wl["encoder"], device=device)
    pairs = [(0.0, r) for r in wl["radii"]]
    model = MultiscaleBoundedFourier(dict(wl["net"]), boundaries=[p for p in pairs for _ in (0, 1)]).to(device)
    optim = FusedAdam(model, lr=LR, betas=(0.9, 0.999), weight_decay=0.0)
    C, H, W = SLICE
    coords, gt, _ = synthetic.make_fit_arrays(1234, C, H, W, image_space=False, normalization="max")
    coords, gt = coords.to(device), gt.to(device)
    dist = torch.sqrt(coords[:, 1] ** 2 + coords[:, 2] ** 2)
    n_fit = (coords.shape[0] // bs) * bs                      # whole batches only: one graph
    trainer = FusedTrainer(model, enc, optim, "LSL", bs, coords[:n_fit], gt[:n_fit], None, wl["loss_opts"], dist=dist[:n_fit],
                           consistency=(pairs, 0.1))

    def run(n):
        for _ in range(n):
            last = trainer.step()
        return last

    run(max(args.warmup, 3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(20)
    torch.cuda.synchronize()
    run(min(int(1.5 / max((time.perf_counter() - t0) / 20, 1e-6)), 2000))      # clocks ramp over ~1.5 s of load
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        torch.cuda.synchronize()
        e0.record()
        last = run(args.steps)
        e1.record()
        clk.sample()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    value = args.steps * bs / (ms * 1e-3)
    loss_last = float(last)
    # end to end: host batches in (pinned staging block -> one H2D copy), the distances are derived from the coordinates on
    # the device, one graph launch per step, every loss read on the host
    eng = trainer.eng
    dbuf = [torch.empty(bs, device=device) for _ in range(2)]
    opts = dict(trainer.loss_opts)

    def host_step(c, y, m, slot, b):
        torch.sqrt(c[:, 1] ** 2 + c[:, 2] ** 2, out=dbuf[slot])
        eng.train_step("LSL", c, y, b, loss_opts=opts, use_cursor=False, dist=dbuf[slot])

    stepper = HostFedStepper(eng, "LSL", bs, masked=False, loss_opts=opts, depth=2, step_fn=host_step)
    h_c, h_g = coords[: bs * 8].cpu(), gt[: bs * 8].cpu()
    losses = []

    def e2e_run(n, i0=0):
        prev = None
        for i in range(i0, i0 + n):
            j = (i % 8) * bs
            t = stepper.submit(h_c[j:j + bs], h_g[j:j + bs])
            if prev is not None:
                losses.append(stepper.loss(prev))
            prev = t
        losses.append(stepper.loss(prev))

    e2e_run(6)
    torch.cuda.synchronize()
    losses.clear()
    t0 = time.perf_counter()
    e2e_run(args.steps, 6)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    # CPU port: the oracle's multiscale forward + the same composite loss + torch autograd + Adam on the host cores
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import inr_oracle as O
        from metrics.losses import ConsistencyLoss, LogSpaceLoss
        lsl, cons = LogSpaceLoss(wl["loss_opts"]), ConsistencyLoss(pairs)
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(1234)
        encB = O.encoder_init(wl["encoder"])
        sd = O.multiscale_init(dict(wl["net"]), bounded=True)
        P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(P.values()), lr=LR)
        rows = 20000
        cc, yy = coords[:rows * 3].cpu(), gt[:rows * 3].cpu()
        ddc = dist[:rows * 3].cpu()
        bounds = [p for p in pairs for _ in (0, 1)]
        n_cpu, t_cpu = 0, 0.0
        for i in range(40):
            if t_cpu > 12.0:
                break
            k = i % 3
            c, y, d = cc[k * rows:(k + 1) * rows], yy[k * rows:(k + 1) * rows], ddc[k * rows:(k + 1) * rows]
            t0 = time.perf_counter()
            outs = O.multiscale_forward(P, O.encode(c, encB, "gauss"), wl["net"]["network_depth"], d, bounds)
            opt.zero_grad()
            loss = 0.1 * cons(outs, d)
            for o in outs:
                loss = loss + 0.5 * lsl(o.contiguous(), y)
            loss.backward()
            opt.step()
            if i > 0:
                n_cpu += 1
                t_cpu += time.perf_counter() - t0
        cpu = {"value": n_cpu * rows / t_cpu, "unit": "coords/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{n_cpu} steps x {rows} coords of the {bs}-coord batch, oracle port (torch CPU fp32, autograd, Adam), "
                         f"{torch.get_num_threads()} threads, {t_cpu:.1f} s"}
    step_tflops = wl["flop_per_coord"] * bs / (ms / args.steps * 1e-3) / 1e12
    n_launch = 1 + 9 + 3 + 3 + 8 + 1 + 1          # encode, 9 stage GEMMs, ms head / scalars / dout, dout amax / scalars / top, 8 dgrad GEMMs, wgrad, Adam
    line = {
        "metric": "train coords/sec (fwd+bwd+Adam)", "value": value, "unit": "coords/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate/master", "data": "synthetic",
        "config": {"workload": wl["name"], "model": wl["model"], "batch_per_gpu": bs, "slice": list(SLICE),
                   "parallelism": "single GPU", "launch": "cuda graph" if trainer.use_graph else "eager",
                   "inputs": f"resident coords+targets+distances {n_fit * 24 / 2**20:.0f} MiB, walked in grid order (device cursor); "
                             "gauss encoding in-kernel",
                   "step": f"{n_launch} kernels: src/train_kspace_multiscale.py loop body fused -- encoding, 9 stage GEMMs (BoundedLinear row "
                           "masks in the epilogues), 4 heads + 4 x LogSpaceLoss + 0.1 ConsistencyLoss, backward entry, 8 dgrad stage GEMMs, "
                           "split-K wgrad, Adam + repack", "loss": "LSL + consistency", "loss_last_step": loss_last},
        "clocks": clk.summary(),
        "e2e": {"value": args.steps * bs / (ms_e2e * 1e-3), "unit": "coords/s", "h2d_bytes_per_step": stepper.h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "api": "trainer.HostFedStepper over ChainEngine.train_step(dist=...) (C ABI inr_train_step_dist): "
                "host batch -> pinned staging -> H2D -> one graph launch per step [distances, kernels, D2H loss], every loss read on the host",
                "loss_last_step": losses[-1]},
        "gpu_launches": n_launch * args.steps,
        "roofline": {"bound": "tensor", "achieved": step_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": step_tflops / peaks["tflops_sustained"], "traffic": None,
                     "kernel": "whole step (live-graph algorithmic FLOP of SURVEY 8d / step time)",
                     "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['source']})"},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + [MULTISCALE_WORKLOAD["name"]])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU context measurement")
    ap.add_argument("--no-secondary", action="store_true", help="default run: skip the SIREN configs[0] measurement")
    ap.add_argument("--parallel", default="dp", choices=["dp", "independent"],
                    help="N>1: dp = one fit, per-GPU batch shard + NCCL gradient all-reduce; independent = one fit per GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == MULTISCALE_WORKLOAD["name"]:
        if args.impl == "reference" or args.gpus != 1:
            raise SystemExit("the multiscale workload is measured on one GPU through the module face only")
        return run_multiscale(args)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)
    line = run_workload(args, args.workload, dist, rank, world, local, device, cpu_baseline=not args.no_cpu_baseline)
    # BASELINE.json's metric is quoted on "SIREN/WIRE": the default single-GPU run also measures configs[0] (SIREN d4 w256,
    # image space, L2, batch 10000) and carries its numbers in the same line
    if rank == 0 and world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_secondary:
        sec = run_workload(args, "siren_image_l2_bs10000", None, 0, 1, local, device, cpu_baseline=False)
        line["workloads"] = {"siren_image_l2_bs10000": {k: sec[k] for k in ("value", "unit", "ms_per_step", "e2e", "roofline",
                                                                             "gpu_launches", "config", "gpu_eager_baseline")}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    # a CUDA graph that captured NCCL kernels can wedge destroy_process_group(); results are out, leave hard
    if dist:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def gpu_eager_steps(wl, device, seconds=2.0):
    """The reference's own arithmetic on THIS GPU ("the real bar"): the oracle port -- plain torch forward,
    autograd backward, explicit Adam, fp32 / complex64, cuBLAS -- with its tensors on the B200, same batch.  Context for
    the headline only; bounded to a couple of seconds."""
    from oracle import inr_oracle as O
    bs = wl["batch"]
    torch.manual_seed(1234)
    encB = O.encoder_init(wl["encoder"])
    sd = {k: v.to(device) for k, v in O.MODEL_INIT[wl["model"]](dict(wl["net"])).items()}
    encB = None if encB is None else encB.to(device)
    n = 6
    coords = torch.rand(bs * n, 3, device=device) * 2 - 1
    gt = (torch.randn(bs * n, 2, device=device) * 0.05) if not wl["image_space"] else torch.rand(bs * n, 2, device=device)
    mask = (torch.arange(bs * n, device=device) % 2 == 0) if wl["undersampling"] else None
    opts = None
    if wl["loss_opts"] and "hdr_eps" in wl["loss_opts"]:
        opts = {"sigma": wl["loss_opts"]["hdr_ff_sigma"], "eps": wl["loss_opts"]["hdr_eps"], "factor": wl["loss_opts"]["hdr_ff_factor"]}
    run = lambda k: O.train_steps(wl["model"], wl["net"], sd, encB, wl["encoder"]["embedding"], coords, gt, k, bs, LR, wl["loss"],
                                  opts, mask=mask)
    run(2)
    torch.cuda.synchronize()
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds and done < 200:
        run(n)
        torch.cuda.synchronize()
        done += n
    dt = time.perf_counter() - t0
    return {"value": done * bs / dt, "unit": "coords/s", "ms_per_step": dt / done * 1e3, "kind": "oracle port in PyTorch eager on the same B200 (fp32 / complex64, autograd, explicit Adam; separable HDR)",
            "sample": f"{done} steps x {bs} coords"}


def run_workload(args, workload, dist, rank, world, local, device, cpu_baseline=True):
    """Measures one workload on this rank's GPU; returns the JSON line (rank 0) or None."""
    wl = WORKLOADS[workload]
    from mri_implicit_neural_representations_b200.parallel import allreduce_mean_
    bs = wl["batch"]
    peaks = load_peaks()

    # ---- state.  dp: identical replicas (same seed), every rank walks its own slice set, gradients all-reduced
    # (SURVEY 8e-2).  independent: one fit per rank, no data-path collective.
    dp = world > 1 and args.parallel == "dp"
    eng, tensors, encB = build_engine(wl, device, seed=1234 if dp else 1234 + rank)
    coords, gt, mask = resident_arrays(wl, device, 1234 + 100 * rank, min_bytes=200 << 20)
    n_rows = coords.shape[0]
    wire = wl["model"] in ("WIRE", "WIRE2D")
    mfn = wl["model"] in ("Gabor", "KGabor", "Fourier")
    steps_per_pass = n_rows // bs
    out_buf = torch.empty(bs, 2, device=device) if (wl["loss_opts"] or {}).get("tv") else None   # the TV term reads the output

    # dp gradient exchange: fused into the optimiser kernel over NVLink peer memory (symmetric memory, double-buffered
    # by step parity); NCCL all-reduce only if symmetric memory cannot be set up on this box
    peer, exchange = None, "none"
    if dp:
        try:
            if os.environ.get("INR_DP_EXCHANGE", "peer") != "peer":
                raise RuntimeError("disabled by INR_DP_EXCHANGE")
            from mri_implicit_neural_representations_b200.parallel import PeerGradExchange
            peer = PeerGradExchange(eng.plan.n_params, device)
            exchange = "optimiser kernel reads all ranks' fp32 gradients over NVLink peer memory (flag barrier, no all-reduce kernel)"
        except Exception as e:
            ok = torch.tensor([0], device=device)
            exchange = f"NCCL all-reduce(avg) ({type(e).__name__}: {e})"[:200]
        if peer is not None:        # all ranks must agree on the mechanism
            ok = torch.tensor([1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0 and peer is not None:
            peer, exchange = None, "NCCL all-reduce(avg) (symmetric memory unavailable on another rank)"
    gstep = [0]                                 # host count of optimiser steps (same on every rank): exchange-buffer parity

    def one_step(c=None, y=None, m=None, use_cursor=True):
        c = coords if c is None else c
        y = gt if y is None else y
        m = mask if c is coords else m
        if dp and peer is not None:
            par = gstep[0] & 1
            eng.grad_step(wl["loss"], c, y, bs, mask=m, loss_opts=wl["loss_opts"], use_cursor=use_cursor, out=out_buf,
                          grads=peer.grads(par))
            eng.adam_step_peers(peer, par)
        elif dp:    # forward+loss+backward -> all-reduce(mean) of the flat fp32 gradients -> Adam (+ fp16 re-pack)
            eng.grad_step(wl["loss"], c, y, bs, mask=m, loss_opts=wl["loss_opts"], use_cursor=use_cursor, out=out_buf)
            allreduce_mean_(eng.grads)
            eng.adam_step()
        else:
            eng.train_step(wl["loss"], c, y, bs, mask=m, loss_opts=wl["loss_opts"], use_cursor=use_cursor, out=out_buf)
        gstep[0] += 1

    # eager warm-up (also sets kernel attributes outside capture), then capture the step in CUDA graphs
    # (one per exchange-buffer parity)
    for _ in range(4):
        one_step()
    torch.cuda.synchronize()
    eng.cursor.zero_()          # the eager steps advanced the device cursor: the host mirror `pos` below starts at 0 as well
    graph_mode = "cuda graph"
    n_graphs = 2 if peer is not None else 1
    try:
        graphs = []
        for par in range(n_graphs):
            gstep[0] = par                      # capture records pointers only; nothing executes here
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                one_step()
            graphs.append(gph)
        gstep[0] = 0                            # 4 eager steps done: parity 0 is next

        class _Replay:
            def replay(self):
                graphs[gstep[0] % n_graphs].replay()
                gstep[0] += 1
        graph = _Replay()
    except Exception as e:          # e.g. a collective that refuses capture: fall back to eager launches
        graph_mode = f"eager ({type(e).__name__})"
        torch.cuda.synchronize()
        gstep[0] = 0

        class _Eager:
            def replay(self):
                one_step()
        graph = _Eager()

    pos = [0]                                   # host mirror of the device-side batch cursor, in steps

    def reset_cursor():
        eng.cursor.zero_()
        pos[0] = 0

    def run_steps(n):
        while n > 0:
            if pos[0] >= steps_per_pass:
                reset_cursor()
            chunk = min(n, steps_per_pass - pos[0])
            for _ in range(chunk):
                graph.replay()
            pos[0] += chunk
            n -= chunk

    # ---- device-resident throughput (`value`)
    # clocks ramp over tens of ms from idle: warm up for at least W steps AND ~1.5 s of continuous load
    run_steps(args.warmup)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_steps(50)
    torch.cuda.synchronize()
    n_ramp = torch.tensor([int(1.5 / max((time.perf_counter() - t0) / 50, 1e-6))], device=device)
    if dist:                                     # every rank must run the same number of steps (collectives / flag epochs)
        dist.all_reduce(n_ramp, op=dist.ReduceOp.MAX)
    run_steps(min(int(n_ramp), 200000))
    torch.cuda.synchronize()
    reset_cursor()
    run_steps(args.warmup % steps_per_pass)      # leave the cursor where a W-step warm-up would
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # NVML initialisation (tens of ms, different on every rank) happens BEFORE the barrier: nothing but the barrier's own
    # exit skew sits between the ranks' synchronisation point and e0.record()
    with ClockSampler(local) as clk:
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        run_steps(args.steps)        # graph replays; one 4-byte memset when the walk wraps around the resident rows
        e1.record()
        clk.sample()                 # the steps are still running: a reading from inside the timed region
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * args.steps * bs / (ms * 1e-3)
    loss_dev = float(eng.loss_out)

    # data-parallel replicas must be bit-identical after the warm-up, ramp and timed steps (checked BEFORE profile_step below,
    # whose instrumented steps update every replica with its local gradients only): min / max over ranks of an fp64 parameter checksum
    dp_check = None
    if dist and dp:
        torch.cuda.synchronize()
        cs = eng.params.double().sum().reshape(1)
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dp_check = {"param_checksum_min": float(lo), "param_checksum_max": float(hi), "identical": float(lo) == float(hi),
                    "steps_taken": int(eng.step)}

    # ---- per-kernel device times (CUDA events between the four kernels of a step, same stream)
    reset_cursor()
    prof = eng.profile_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"], reps=100 if bs <= 25000 else 20,
                            out=out_buf)
    # WIRE: first, ONE chained launch of the depth forward GEMMs, last (+ step scalars in its last CTA), blast, ONE chained
    # launch of the depth dgrad GEMMs, wgrad, Adam = 7; WIRE2D the same plus its separate scalars kernel = 8
    wide = bool(getattr(eng.plan, "wide", False))
    n_sine = max(wl["net"]["network_depth"] - 1, 1)
    n_launch = (6 if wl["model"] == "WIRE" else 8) if wire else (4 if not mfn else (36 if wl["model"] != "Fourier" else 23))
    if wide:          # encoding, n_sine layer GEMMs, head + loss, scalars, backward entry, n_sine - 1 dgrad GEMMs, wgrad, Adam
        n_launch = 1 + n_sine + 1 + 1 + 1 + (n_sine - 1) + 1 + 1
    # roofline kernel.  WIRE: the chained forward GEMM launch (all hidden layers in one persistent launch), timed by the
    # events around it; WIRE2D: the average of its per-layer forward GEMM launches; SIREN / FFN: the fused forward kernel
    chained = wire
    n_gemm_launches = 1 if chained else wl["net"]["network_depth"]
    kern_ms = prof["forward_layer_gemms"] / n_gemm_launches if wire else prof["forward"]
    kern_flop = (wl["fwd_flop_per_coord"] / n_gemm_launches if wire else wl["fwd_flop_per_coord"]) * bs
    fwd_tflops = kern_flop / (kern_ms * 1e-3) / 1e12
    step_tflops = wl["flop_per_coord"] * bs / (ms / args.steps * 1e-3) / 1e12

    # ---- end to end through the public API with HOST buffers: every step copies its batch host -> pinned staging -> HBM,
    # runs the fused step and reads the step's loss back to the host.  HostFedStepper makes that ONE graph launch per step
    # (one H2D block copy on a copy stream + kernels + D2H loss) over two rotating staging sets, so the host prepares step i+1 while step i runs;
    # every loss is read (one step late) inside the timed region.
    from mri_implicit_neural_representations_b200.trainer import HostFedStepper
    h_coords = coords[: bs * 64].cpu()
    h_gt = gt[: bs * 64].cpu()
    h_mask = None if mask is None else mask[: bs * 64].cpu()
    par_off = [0]

    def host_step(c, y, m, slot, b):
        gstep[0] = slot ^ par_off[0]            # exchange-buffer parity follows the staging slot (both alternate per step)
        one_step(c, y, m, use_cursor=False)

    stepper = HostFedStepper(eng, wl["loss"], bs, masked=mask is not None, loss_opts=wl["loss_opts"], depth=2,
                             step_fn=host_step, out=out_buf, use_graph=graph_mode == "cuda graph")
    par_off[0] = gstep[0] & 1
    e2e_losses = []

    def e2e_run(n, i0=0):
        prev = None
        for i in range(i0, i0 + n):
            j = (i % 64) * bs
            t = stepper.submit(h_coords[j:j + bs], h_gt[j:j + bs], None if h_mask is None else h_mask[j:j + bs])
            if prev is not None:
                e2e_losses.append(stepper.loss(prev))
            prev = t
        e2e_losses.append(stepper.loss(prev))

    n_e2e_warm = 2 * max(5, args.warmup // 4)    # even: staging-slot (and exchange-buffer) parity carries over
    e2e_run(n_e2e_warm)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    n_e2e = args.steps
    e2e_losses.clear()
    t0 = time.perf_counter()
    e0.record()
    e2e_run(n_e2e, n_e2e_warm)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    assert len(e2e_losses) == n_e2e
    ms_e2e = max(e0.elapsed_time(e1), wall * 1e3)
    if dist:
        t = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    e2e_value = world * n_e2e * bs / (ms_e2e * 1e-3)

    if rank != 0:
        return None

    eager = None
    if world == 1 and not args.no_gpu_eager and wl["model"] in ("WIRE", "SIREN", "FFN"):
        try:
            eager = gpu_eager_steps(wl, device)
        except Exception as e:      # context only: never fail the bench over it
            eager = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    cpu = None
    if cpu_baseline and world == 1:
        v1, dt1 = cpu_port_steps(wl, 1, 1, bs)
        n_cpu = max(2, min(200, int(12.0 / max(dt1, 1e-3))))
        v, dt = cpu_port_steps(wl, n_cpu, 1, bs)
        cpu = {"value": v, "unit": "coords/s", "cores": os.cpu_count() or 1, "kind": "port",
               "port": "oracle restatement of the reference modules (separable HDR: faster than the reference's [bs, m] outer product)",
               "sample": f"{n_cpu} steps x {bs} coords, oracle port (torch CPU fp32, autograd, Adam), {torch.get_num_threads()} threads, {dt:.1f} s"}

    clocks = clk.summary()
    line = {
        "metric": "train coords/sec (fwd+bwd+Adam)", "value": value, "unit": "coords/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate/master", "data": "synthetic",
        "config": {"workload": workload, "model": wl["model"], "batch_per_gpu": bs, "slice": list(SLICE),
                   "parallelism": ("single GPU" if world == 1 else
                                   (f"dp{world}: replicated weights, per-GPU batch {bs}, {eng.plan.n_params * 4} B fp32 gradients per step; exchange: {exchange}"
                                    if dp else f"independent fit per GPU x{world}, no collective")),
                   "launch": graph_mode,
                   "inputs": f"resident coords+targets {n_rows * 20 / 2**20:.0f} MiB > 126 MB L2, walked in grid order (cold each step)",
                   "step": (f"{n_launch} kernels: first layer, {wl['net']['network_depth']} layer GEMMs (3-pass split fp16 + complex Gabor epilogue"
                            "; one chained persistent launch, tiles handed from layer to layer), "
                            f"final layer + loss{' + step scalars (last CTA)' if wl['model'] == 'WIRE' else ', scalars'}, "
                            f"{'final-layer backward as the first items of the dgrad chain + ' if wl['model'] == 'WIRE' else 'final-layer backward, '}"
                            f"{wl['net']['network_depth']} dgrad layer GEMMs (one chained launch), split-K wgrad, "
                            "complex Adam + repack") if wire else
                           (f"{n_launch} kernels: encoding, |mu|^2, 9 x (envelope GEMM + stage GEMM), head + loss, TV, scalars, top stage, "
                            "8 dgrad stage GEMMs, split-K wgrad, d mu / d gamma, Adam + repack") if mfn else
                           (f"{n_launch} kernels: encoding, {n_sine} streamed layer GEMMs (sine epilogue), head + loss, scalars, backward entry, "
                            f"{n_sine - 1} dgrad layer GEMMs, split-K wgrad, Adam + repack (wide chain: width {wl['net']['network_width']})") if wide else
                           "4 kernels: fused forward+loss, dgrad chain, split-K wgrad, Adam+repack",
                   "loss": wl["loss"], "undersampling": wl["undersampling"],
                   "loss_last_step": loss_dev},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "coords/s", "h2d_bytes_per_step": stepper.h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / n_e2e, "api": "trainer.HostFedStepper.submit / .loss over ChainEngine.train_step (C ABI inr_train_step): host batch -> pinned staging block -> H2D on a copy stream -> one graph launch per step [kernels, D2H loss], two staging sets, every loss read on the host",
                "loss_last_step": e2e_losses[-1]},
        "gpu_launches": (n_launch + (1 if dp else 0)) * args.steps,     # dp: gradient-reduce kernel + optimiser kernel instead of one
        "roofline": {"bound": "tensor", "achieved": fwd_tflops, "peak": peaks["tflops_burst"], "unit": "TFLOP/s",
                     "frac": fwd_tflops / peaks["tflops_burst"],
                     # DRAM read+write bytes of ONE launch of this kernel from the committed `ncu --set full` capture
                     "traffic": load_traffic(workload)[0], "traffic_source": load_traffic(workload)[1],
                     "kernel": (f"lgemm_kernel ({wl['model']} forward: the {wl['net']['network_depth']} hidden-layer GEMMs + Gabor epilogues as one chained persistent launch)" if wire else
                                ("forward phase (encoding + 18 lgemm launches + head/loss + TV)" if mfn else
                                 ("forward phase (encoding + layer GEMMs + head/loss)" if wide else "chain_fwd_kernel<SIN>"))),
                     "kernel_ms": kern_ms,
                     "issued_tflops": (wl["issued_fwd_flop_per_coord"] / n_gemm_launches * bs / (kern_ms * 1e-3) / 1e12) if wire else None,
                     "peak_source": f"MEASURED_PEAKS.json bf16 burst ({peaks['source']})",
                     "step_tflops": step_tflops, "step_frac_of_sustained": step_tflops / peaks["tflops_sustained"],
                     "kernels_ms": prof},
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
    }
    if dp_check is not None:
        line["dp_check"] = dp_check
    return line


if __name__ == "__main__":
    main()
