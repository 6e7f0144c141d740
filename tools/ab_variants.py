"""In-process A/B of kernel variants selected by environment variables that the library reads PER LAUNCH: the variants are
timed round-robin (inr_profile_step, event-timed kernels of one eager step) so that clock and thermal drift hit all alike.
usage: python tools/ab_variants.py <workload> "<NAME=VAL,...>" "<NAME=VAL,...>" ...   (an empty string = defaults)"""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1]
variants = sys.argv[2:]
wl = bench.WORKLOADS[name]
bs = wl["batch"]
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.randn(bs, 2, device=dev) * 0.05
mask = (torch.arange(bs, device=dev) % 2 == 0).to(torch.uint8) if wl["undersampling"] else None
t_end = time.time() + 2.0
while time.time() < t_end:
    for _ in range(50):
        eng.train_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"])
    torch.cuda.synchronize()
keys = sorted({kv.split("=")[0] for v in variants for kv in v.split(",") if kv})
res = {v: [] for v in variants}
for rnd in range(int(os.environ.get("AB_ROUNDS", "7"))):
    for v in variants:
        for k in keys:
            os.environ.pop(k, None)
        for kv in v.split(","):
            if kv:
                k, val = kv.split("=")
                os.environ[k] = val
        for _ in range(5):
            eng.train_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"])
        res[v].append(eng.profile_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"], reps=10))
for v in variants:
    med = {k: round(statistics.median(r[k] for r in res[v]) * 1e3, 1) for k in res[v][0]}
    print(json.dumps({"variant": v or "(defaults)", **med}))
