"""Small fused training steps of every kernel family, as a target for compute-sanitizer (tools/sanitize.sh): WIRE with the
chained CTA-pair layer GEMMs (hand-over counters, cluster barriers, last-CTA scalar reduction), SIREN chain kernels, a wide
SIREN chain, the fused multi-scale step, split-K wgrad and the optimiser kernels.  Batches are a few thousand rows so that a
run under the sanitizer (10-100x slower) stays well inside the kernels' 4 s wait watchdogs.
usage: python tools/sanitize_target.py [wire|siren|wide|multiscale|all] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mri_implicit_neural_representations_b200 as inr
from mri_implicit_neural_representations_b200 import init as pinit

which = sys.argv[1] if len(sys.argv) > 1 else "all"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
ENC = {"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3}


def run(name, model, net, enc, loss, bs, opts=None, masked=False, dist=False):
    torch.manual_seed(1)
    plan = inr.Plan(model, net, enc)
    eng = inr.ChainEngine(plan, max_batch=bs, device=dev, lr=5e-4)
    if model in ("WIRE",):
        tensors = [t for _, t in pinit.wire_tensors(net)]
    elif model in ("BoundedFourier",):
        from oracle import inr_oracle as O      # initial values only (test tool)
        tensors = list(O.multiscale_init({k: v for k, v in net.items() if k != "boundaries"}, bounded=True).values())
    else:
        tensors = [t for _, t in pinit.chain_tensors(model, net)]
    eng.load_tensors(tensors)
    eng.set_encoder(pinit.encoder_matrix(enc))
    c = torch.rand(bs, 3, device=dev) * 2 - 1
    y = torch.randn(bs, 2, device=dev) * 0.05
    m = (torch.arange(bs, device=dev) % 2 == 0).to(torch.uint8) if masked else None
    d = torch.sqrt(c[:, 1] ** 2 + c[:, 2] ** 2) if dist else None
    for _ in range(steps):
        eng.train_step(loss, c, y, bs, mask=m, loss_opts=opts, dist=d)
    torch.cuda.synchronize()
    print(f"{name}: {steps} steps ok, loss {float(eng.loss_out):.5f}", flush=True)


if which in ("wire", "all"):
    run("wire_hdr_masked", "WIRE", {"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256,
                                    "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}, {"embedding": "none"}, "HDR", 2432 + 128,
        {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}, masked=True)      # 20 tiles -> 10 CTA pairs
    run("wire_l2_odd_tiles", "WIRE", {"network_input_size": 3, "network_output_size": 2, "network_depth": 2, "network_width": 256,
                                      "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}, {"embedding": "none"}, "L2", 1100)   # 9 tiles: phantom tile
if which in ("siren", "all"):
    run("siren_l2", "SIREN", {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256}, ENC, "L2", 1300)
if which in ("wide", "all"):
    run("siren_w512_tanh", "SIREN", {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 512,
                                     "last_tanh": True}, ENC, "tanh", 700)
if which in ("multiscale", "all"):
    pairs = [(0.0, 0.4), (0.0, 0.8), (0.0, 1.1), (0.0, 5.0)]
    run("bounded_fourier_lsl", "BoundedFourier", {"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 256,
                                                   "boundaries": [p for p in pairs for _ in (0, 1)]}, ENC, "LSL", 600,
        {"hdr_eps": 1e-2, "consistency": (pairs, 0.1)}, dist=True)
