"""Small fixed workload for ncu: eager fused training steps of BASELINE config 2 (WIRE + HDR + mask, bs 25000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
wl = bench.WORKLOADS["wire_kspace_hdr_bs25000"]
bs = wl["batch"]
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.randn(bs, 2, device=dev) * 0.05
mask = (torch.arange(bs, device=dev) % 2 == 0).to(torch.uint8)
for _ in range(steps):
    eng.train_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"])
torch.cuda.synchronize()
print("loss", float(eng.loss_out))
