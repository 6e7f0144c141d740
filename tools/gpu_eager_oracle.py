"""Context measurement ("also time reference-PyTorch-eager on the B200"): the oracle port -- plain torch
modules-equivalent forward, autograd backward, explicit Adam -- run with its tensors on the GPU, same workloads as
bench.py.  Diagnostic only (not part of the product or of bench.py's reported numbers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import inr_oracle as O

dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["wire_kspace_hdr_bs25000", "siren_image_l2_bs10000"]:
    wl = bench.WORKLOADS[name]
    bs = wl["batch"]
    torch.manual_seed(1234)
    encB = O.encoder_init(wl["encoder"])
    sd = O.MODEL_INIT[wl["model"]](dict(wl["net"]))
    sd = {k: v.to(dev) for k, v in sd.items()}
    encB = None if encB is None else encB.to(dev)
    n = 12
    coords = (torch.rand(bs * n, 3, device=dev) * 2 - 1)
    gt = torch.randn(bs * n, 2, device=dev) * 0.05 if not wl["image_space"] else torch.rand(bs * n, 2, device=dev)
    mask = None
    if wl["undersampling"]:
        mask = (torch.arange(bs * n, device=dev) % 2 == 0)
    opts = None
    if wl["loss_opts"] and "hdr_eps" in wl["loss_opts"]:
        opts = {"sigma": wl["loss_opts"]["hdr_ff_sigma"], "eps": wl["loss_opts"]["hdr_eps"], "factor": wl["loss_opts"]["hdr_ff_factor"]}
    O.train_steps(wl["model"], wl["net"], sd, encB, wl["encoder"]["embedding"], coords, gt, 2, bs, bench.LR, wl["loss"], opts, mask=mask)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    O.train_steps(wl["model"], wl["net"], sd, encB, wl["encoder"]["embedding"], coords, gt, n - 2, bs, bench.LR, wl["loss"], opts,
                  mask=None if mask is None else mask)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: torch eager on GPU (oracle port, fp32 / complex64): {(n - 2) * bs / dt:.3e} coords/s, {dt / (n - 2) * 1e3:.2f} ms/step")
