"""Small fixed workload for ncu: eager fused training steps of BASELINE config 1 (SIREN, bs 10000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
wl = dict(bench.WORKLOADS["siren_image_l2_bs10000"], batch=bs)
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.rand(bs, 2, device=dev)
for _ in range(steps):
    eng.train_step(wl["loss"], coords, gt, bs)
torch.cuda.synchronize()
print("loss", float(eng.loss_out))
