"""WIRE first-contact diagnostics on a B200: prints error numbers for every stage, asserts nothing."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mri_implicit_neural_representations_b200 as inr
from oracle import inr_oracle as O
from oracle import golden_util as G
from oracle.cases import case_setup, loss_and_grad


def rel(a, b):
    a, b = a.cpu(), b.cpu()
    a = torch.view_as_real(a.to(torch.complex128)) if a.is_complex() else a.double()
    b = torch.view_as_real(b.to(torch.complex128)) if b.is_complex() else b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def to64(sd):
    return {k: (v.to(torch.complex128) if v.is_complex() else v.double()) for k, v in sd.items()}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "wire_l2"
    print(torch.cuda.get_device_name(0), name)
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    depth, bs, C = net["network_depth"], coords.shape[0], 181
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    cd, gd = coords.cuda(), gt.cuda()
    md = None if mask is None else mask.to(torch.uint8).cuda()

    print("\n===== forward")
    tr32, tr64 = [], []
    out32 = O.wire_forward(sd, coords, depth, trace=tr32)
    sd64 = to64(sd)
    out64 = O.wire_forward(sd64, coords.double(), depth, trace=tr64)
    out = eng.forward(cd, train=True)
    torch.cuda.synchronize()
    print("out: engine vs fp64", rel(out, out64), "| fp32 oracle vs fp64", rel(out32, out64), "| engine vs fp32", rel(out, out32))
    for l in range(depth + 1):
        h = eng.read_wire_image("h", l + 1, bs)[:bs, :C]
        print(f"h{l+1}: engine vs fp64 {rel(h, tr64[l][1]):.3e} | fp32 oracle vs fp64 {rel(tr32[l][1], tr64[l][1]):.3e}")
        if l >= 1:   # teacher-forced: this layer evaluated in fp64 on the engine's own input
            hin = eng.read_wire_image("h", l, bs)[:bs, :C].cpu().to(torch.complex128)
            z = hin @ sd64[f"net.{l}.linear.weight"].t() + sd64[f"net.{l}.linear.bias"]
            y = O.gabor_act(z, sd64[f"net.{l}.omega_0"], sd64[f"net.{l}.scale_0"])
            print(f"    teacher-forced layer {l}: {rel(h, y):.3e}; ab image {rel(eng.read_wire_image('ab', l, bs)[:bs, :C], z):.3e}")

    print("\n===== fused grad step vs fp64 autograd")
    P = {k: v.clone().requires_grad_(not (k.endswith('omega_0') or k.endswith('scale_0'))) for k, v in sd64.items()}
    o64 = O.wire_forward(P, coords.double(), depth)
    sel, g_sel = (o64, gt.double()) if mask is None else (o64[mask], gt.double()[mask])
    val, dsel = loss_and_grad(loss_kind, opts, sel.detach(), g_sel, coords.double())
    live = [k for k in P if P[k].requires_grad]
    gr64 = dict(zip(live, torch.autograd.grad(sel, [P[k] for k in live], grad_outputs=dsel)))
    out_dev = torch.zeros(bs, 2, device="cuda")
    g = eng.grad_step(loss_kind, cd, gd, bs, mask=md, loss_opts=opts, out=out_dev)
    torch.cuda.synchronize()
    print("loss engine", float(eng.loss_out), "fp64", float(val), "scalars", eng.scalars(bs)[:8].tolist())
    scal = eng.scalars(bs)
    Sl = scal[16:16 + depth + 1].tolist()
    print("per-layer scales", Sl)
    # teacher-forced backward per layer: dZ_{l-1} recomputed in fp64 from the engine's own dZ_l, h_l and (a,b)_{l-1}
    for l in range(depth, 0, -1):
        dz_l = eng.read_wire_image("dz", l, bs)[:bs, :C].cpu().to(torch.complex128) / Sl[l]
        dh = dz_l @ sd64[f"net.{l}.linear.weight"].conj()
        y = eng.read_wire_image("h", l, bs)[:bs, :C].cpu().to(torch.complex128)
        ab = eng.read_wire_image("ab", l - 1, bs)[:bs, :C].cpu().to(torch.complex128)
        w_, s2 = float(sd[f"net.{l-1}.omega_0"]), float(sd[f"net.{l-1}.scale_0"]) ** 2
        Pq = dh.conj() * y
        P, Q = Pq.real, Pq.imag
        da = -2 * s2 * ab.real * P - w_ * Q
        db = -(w_ + 2 * s2 * ab.imag) * P if l - 1 > 0 else torch.zeros_like(da)
        ref = torch.complex(da, db)
        got = eng.read_wire_image("dz", l - 1, bs)[:bs, :C].cpu().to(torch.complex128) / Sl[l - 1]
        print(f"    teacher-forced dZ{l-1} from dZ{l}: {rel(got, ref):.3e}  (amax scaled {float(torch.view_as_real(got).abs().max() * Sl[l-1]):.3e})")
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for k in live:
        print(f"grad {k}: rel err {rel(gv[k], gr64[k]):.3e} (norm {float(torch.view_as_real(gr64[k]).norm() if gr64[k].is_complex() else gr64[k].norm()):.3e})")

    print("\n===== fused train steps vs golden (reference fp32) / fp64 golden")
    eng2 = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng2.load_tensors(list(sd.values()))
    gold = G.load_golden(name)
    losses = []
    for step in range(G.N_ADAM_STEPS):
        eng2.train_step(loss_kind, cd, gd, bs, mask=md, loss_opts=opts)
        torch.cuda.synchronize()
        losses.append(float(eng2.loss_out))
    print("losses engine      ", losses)
    print("losses golden fp32 ", gold["losses"])
    print("losses golden fp64 ", gold["fp64"]["losses"])

    print("\n===== timing")
    for bsx in (25000,):
        engx = inr.ChainEngine(plan, max_batch=bsx, lr=G.LR)
        engx.load_tensors(list(sd.values()))
        cx = torch.rand(bsx, 3, device="cuda") * 2 - 1
        gx = torch.randn(bsx, 2, device="cuda") * 0.05
        mx = (torch.arange(bsx, device="cuda") % 2 == 0).to(torch.uint8)
        for _ in range(30):
            engx.train_step(loss_kind, cx, gx, bsx, mask=mx, loss_opts=opts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            engx.train_step(loss_kind, cx, gx, bsx, mask=mx, loss_opts=opts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        print(f"train_step bs={bsx}: {ms*1000:.1f} us/step -> {bsx/ms*1000:.3e} coords/s")
        print("per-phase ms (fwd, -, bwd+wgrad, adam):", engx.profile_step(loss_kind, cx, gx, bsx, mask=mx, loss_opts=opts, reps=20))


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
