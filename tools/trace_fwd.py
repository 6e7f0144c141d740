"""Print the phase timeline of CTA 0 of the forward kernel (debug trace), bs given on the command line."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mri_implicit_neural_representations_b200 as inr
from mri_implicit_neural_representations_b200 import _lib as L

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
wl = dict(bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "siren_image_l2_bs10000"], batch=bs)
dev = torch.device("cuda", 0)
eng, _, _ = bench.build_engine(wl, dev, 1234)
coords = torch.rand(bs, 3, device=dev) * 2 - 1
gt = torch.rand(bs, 2, device=dev)
for _ in range(200):
    eng.train_step(wl["loss"], coords, gt, bs)
torch.cuda.synchronize()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
L.lib.inr_debug_set_trace(C.c_void_p(buf.data_ptr()))
for rep in range(3):
    buf.zero_()
    eng.train_step(wl["loss"], coords, gt, bs)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    t0 = t[0]
    names = {0: "kernel start", 1: "prologue done", 38: "compute done", 39: "after final sync"}
    for c in range(8): names[2 + c] = f"enc chunk {c} written"
    for l in range(3):
        names[12 + 5 * l] = f"acc L{l} ready (compute)"
        for g in range(4): names[13 + 5 * l + g] = f"epilogue L{l} step {g} done"
        names[40 + l] = f"MMA L{l} issued (mma thread)"
    for i in range(8): names[48 + i] = f"W stage {i} landed (mma thread)"
    ev = sorted((v - t0, names.get(i, str(i))) for i, v in enumerate(t) if v)
    print(f"--- rep {rep} (ns since CTA0 start)")
    prev = 0
    for dt, n in ev:
        print(f"{dt:8d}  (+{dt - prev:6d})  {n}")
        prev = dt
L.lib.inr_debug_set_trace(None)
