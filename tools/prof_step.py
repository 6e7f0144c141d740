"""Event-timed phases of one training step (inr_profile_step) for a bench workload; timing experiments via INR_LGEMM_DBG.
usage: python tools/prof_step.py [workload] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
wl = bench.WORKLOADS[name]
bs = wl["batch"]
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.randn(bs, 2, device=dev) * 0.05
mask = (torch.arange(bs, device=dev) % 2 == 0).to(torch.uint8) if wl["undersampling"] else None
import time
t_end = time.time() + float(os.environ.get("INR_WARM_S", "2.0"))      # SM clocks ramp over ~1.5 s of continuous load
while time.time() < t_end:
    for _ in range(50):
        eng.train_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"])
    torch.cuda.synchronize()
ms = eng.profile_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"], reps=reps)
print(json.dumps({"workload": name, "dbg": os.environ.get("INR_LGEMM_DBG", "0"), **{k: round(v * 1e3, 1) for k, v in ms.items()}}))
