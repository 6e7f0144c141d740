"""Per-CTA phase timeline of one layer-GEMM launch of a training step (debug trace).
usage: INR_TRACE_LGEMM=<launch index within the step> python tools/trace_lgemm.py [workload]"""
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


import time
t_end = time.time() + float(os.environ.get("INR_TRACE_WARM_S", "2.0"))      # SM clocks ramp over ~1.5 s of continuous load
while time.time() < t_end:
    for _ in range(50):
        step_fn()
    torch.cuda.synchronize()
buf = torch.zeros(64 + 32 * 1024, dtype=torch.int64, device=dev)
for sel in os.environ.get("INR_TRACE_LGEMMS", os.environ.get("INR_TRACE_LGEMM", "0")).split(","):
    os.environ["INR_TRACE_LGEMM"] = sel
    buf.zero_()
    L.lib.inr_debug_set_trace(C.c_void_p(buf.data_ptr()))
    step_fn()
    torch.cuda.synchronize()
    L.lib.inr_debug_set_trace(None)
    t = buf.cpu()[64:].view(-1, 32)
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    print(f"--- lgemm launch {sel}: {t.shape[0]} CTAs; ns since first CTA start: start, prologue | per item: mma issued, acc ready, epilogue done | exit")
    for i in list(range(0, t.shape[0], max(1, t.shape[0] // 8))) + [1, 75]:
        v = t[i].tolist()
        f = lambda x: f"{int(x) - t0:7d}" if x else "      -"
        items = " | ".join(" ".join(f(v[2 + 3 * j + k]) for k in range(3)) for j in range(4))
        print(f"cta {i:4d} {f(v[0])} {f(v[1])} | {items} | {f(v[15])}")
        if any(v[16:24]):
            print(f"          producer cycles: flag wait {v[16]}, empty wait {v[17]}, copy issue {v[18]} | mma thread cycles: acc wait {v[20]}, "
                  f"full wait {v[21]}, issue+commit {v[22]} over {v[23]} items")
    print("max exit", int(t[:, 15].max()) - t0)
    v = t[t.shape[0] // 2].tolist()
    if v[25]:
        print(f"producer warp of cta {t.shape[0] // 2}: {v[24]} cycles in {v[25]} ns -> SM clock {v[24] / v[25] * 1e3:.0f} MHz")
