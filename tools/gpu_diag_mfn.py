"""MFN first-contact diagnostics on a B200 (FourierNet / MultiscaleKFourier): prints error numbers, asserts nothing."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mri_implicit_neural_representations_b200 as inr
from mri_implicit_neural_representations_b200 import init as pinit
from oracle import inr_oracle as O
from oracle import golden_util as G
from oracle.cases import case_setup, loss_and_grad


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def main():
    name = "fourier_l2"
    print(torch.cuda.get_device_name(0), name)
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    L, bs = net["network_depth"], coords.shape[0]
    plan = inr.Plan("Fourier", net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    cd, gd = coords.cuda(), gt.cuda()
    x = O.encode(coords, encB, "gauss")
    tr = []
    out_ref = O.mfn_forward(sd, x, L, False, trace=tr)
    print("\n===== forward")
    out = eng.forward(cd, train=True)
    torch.cuda.synchronize()
    print("out rel err", rel(out, out_ref))
    for i in range(L + 1):
        print(f"z{i} rel err {rel(eng.read_mfn_image('z', i, bs)[:bs], tr[i]):.3e}")
    print("\n===== fused grad step vs fp32 autograd")
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.mfn_forward(P, x, L, False)
    val, dout = loss_and_grad(loss_kind, opts, o.detach(), gt, coords)
    gr = dict(zip(P.keys(), torch.autograd.grad(o, list(P.values()), grad_outputs=dout)))
    g = eng.grad_step(loss_kind, cd, gd, bs, loss_opts=opts)
    torch.cuda.synchronize()
    sc = eng.scalars(bs)
    print("loss engine", float(eng.loss_out), "oracle", float(val), "S", float(sc[1]), "stage scales", sc[16:16 + L + 1].tolist())
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for k in sd:
        print(f"grad {k}: rel err {rel(gv[k], gr[k]):.3e} (norm {float(gr[k].norm()):.3e})")
    print("\n===== fused train steps vs golden")
    eng2 = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng2.load_tensors(list(sd.values())); eng2.set_encoder(encB)
    gold = G.load_golden(name)
    losses = []
    for _ in range(G.N_ADAM_STEPS):
        eng2.train_step(loss_kind, cd, gd, bs, loss_opts=opts)
        torch.cuda.synchronize()
        losses.append(float(eng2.loss_out))
    print("losses engine", losses)
    print("losses golden", gold["losses"])
    for (off, rows, cols, layer, is_bias), k in list(zip(plan.tensors, sd.keys()))[:6]:
        dg = G.tensor_digest(eng2.params[off:off + rows * cols].cpu())
        print(f"final {k}: l2 {dg['l2']:.8e} vs golden {gold['final'][k]['l2']:.8e}")

    print("\n===== multiscale (autograd face: 4 heads, external dL/dout)")
    torch.manual_seed(5)
    msd = O.multiscale_init(dict(net), bounded=False)
    mplan = inr.Plan("MultiscaleFourier", net, enc_cfg)
    meng = inr.ChainEngine(mplan, max_batch=bs, lr=G.LR)
    meng.load_tensors(list(msd.values())); meng.set_encoder(encB)
    PM = {k: v.clone().requires_grad_(True) for k, v in msd.items()}
    outs = O.multiscale_forward(PM, x, L)
    mout = meng.forward(cd, train=True)
    torch.cuda.synchronize()
    for k, oo in enumerate(outs):
        print(f"head {k} out rel err {rel(mout[:, 2 * k:2 * k + 2], oo):.3e}")
    douts = [torch.randn(bs, 2) * 1e-3 for _ in outs]
    live = [k for k in PM]
    grs = torch.autograd.grad(outs, [PM[k] for k in live], grad_outputs=douts, allow_unused=True)
    mg = meng.backward(torch.cat(douts, dim=1).cuda())
    torch.cuda.synchronize()
    mgv = dict(zip(msd.keys(), meng._views(meng.grads)))
    for k, gr_ in zip(live, grs):
        if gr_ is None:
            print(f"grad {k}: reference grad None; engine norm {float(mgv[k].norm()):.3e}")
        elif "linear.0" in k or "linear.6" in k or "filters.0" in k or "filters.7" in k or "output_linear.3" in k or "output_linear.7" in k:
            print(f"grad {k}: rel err {rel(mgv[k], gr_):.3e}")

    print("\n===== timing (Fourier d8 w512, L2)")
    for bsx in (30000, 100000):
        engx = inr.ChainEngine(plan, max_batch=bsx, lr=G.LR)
        engx.load_tensors(list(sd.values())); engx.set_encoder(encB)
        cx = torch.rand(bsx, 3, device="cuda") * 2 - 1
        gx = torch.rand(bsx, 2, device="cuda")
        for _ in range(10):
            engx.train_step(loss_kind, cx, gx, bsx, loss_opts=opts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            engx.train_step(loss_kind, cx, gx, bsx, loss_opts=opts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print(f"train_step bs={bsx}: {ms*1000:.1f} us/step -> {bsx/ms*1000:.3e} coords/s ({22026240*bsx/ms/1e9:.1f} TFLOP/s algorithmic)")
        print("phases ms:", engx.profile_step(loss_kind, cx, gx, bsx, loss_opts=opts, reps=10))
        del engx


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
