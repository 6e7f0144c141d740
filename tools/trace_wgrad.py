"""Per-CTA phase timeline of the split-K wgrad kernel (debug trace) for a bench workload."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mri_implicit_neural_representations_b200 import _lib as L

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
wl = bench.WORKLOADS[name]
bs = wl["batch"]
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.randn(bs, 2, device=dev) * 0.05
mask = (torch.arange(bs, device=dev) % 2 == 0).to(torch.uint8) if wl["undersampling"] else None


def step_fn():
    eng.train_step(wl["loss"], coords, gt, bs, mask=mask, loss_opts=wl["loss_opts"])


for _ in range(50):
    step_fn()
torch.cuda.synchronize()
buf = torch.zeros(64 + 8 * 1024, dtype=torch.int64, device=dev)
L.lib.inr_debug_set_trace(C.c_void_p(buf.data_ptr()))
for rep in range(2):
    buf.zero_()
    step_fn()
    torch.cuda.synchronize()
    t = buf.cpu()[64:].view(-1, 8)
    live = t[:, 0] > 0
    t = t[live]
    t0 = int(t[:, 0].min())
    print(f"--- rep {rep}: {t.shape[0]} CTAs; columns = ns since first CTA start: start, prologue, first operands, mma issued, acc ready, epilogue done, exit")
    for i in list(range(0, t.shape[0], max(1, t.shape[0] // 24))):
        print(f"cta {i:4d} " + " ".join(f"{int(v) - t0:8d}" if v else "       -" for v in t[i, :7].tolist()))
    print("max exit", int(t[:, 6].max()) - t0, " median mma-issued->acc", int((t[:, 4] - t[:, 3]).median()), " median epilogue", int((t[:, 5] - t[:, 4]).median()))
L.lib.inr_debug_set_trace(None)
