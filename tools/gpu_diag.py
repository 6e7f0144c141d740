"""First-contact diagnostics on a B200: prints error numbers for every stage, asserts nothing."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mri_implicit_neural_representations_b200 as inr
from oracle import inr_oracle as O
from oracle import golden_util as G
from oracle.cases import case_setup, loss_and_grad


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def stage(name):
    print(f"\n===== {name}", flush=True)


def main():
    print(torch.cuda.get_device_name(0))
    stage("selftest")
    for mode in (0, 1, 2):
        for variant in (0,):   # variant 1 (LBO/SBO swapped) faults with an illegal address: convention confirmed
            try:
                err, ref = inr.selftest_umma(mode, variant)
                print(f"mode {mode} variant {variant}: max_abs_err {err:.4e} ref_absmax {ref:.4e}", flush=True)
            except Exception as e:
                print(f"mode {mode} variant {variant}: FAILED {e}", flush=True)
                return

    name = sys.argv[1] if len(sys.argv) > 1 else "siren_l2"
    stage(f"case {name}: forward")
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup(name)
    plan = inr.Plan(model_kind, net, enc_cfg)
    bs = coords.shape[0]
    eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng.load_tensors(list(sd.values()))
    eng.set_encoder(encB)
    depth = net["network_depth"]
    x = O.encode(coords, encB, enc_cfg["embedding"])
    tr = []
    out_ref = O.model_forward(model_kind, sd, x, net, trace=tr)
    cd, gd = coords.cuda(), gt.cuda()
    out = eng.forward(cd, train=True)
    torch.cuda.synchronize()
    print("out rel err", rel(out, out_ref), "| ref norm", float(out_ref.norm()), "engine norm", float(out.norm()))
    print("H0 (encoding) rel err", rel(eng.read_image("h", 0, bs)[:bs], x))
    for l in range(depth - 1):
        z, h = tr[l]
        print(f"H{l+1} rel err", rel(eng.read_image("h", l + 1, bs)[:bs], h))
        if model_kind == "SIREN":
            dref = torch.cos(30.0 * z)
        else:
            dref = (z > 0).float()
        print(f"D{l} rel err", rel(eng.read_image("d", l, bs)[:bs], dref))

    stage("backward (external dout)")
    val, dout = loss_and_grad(loss_kind, opts, out_ref, gt, coords)
    if model_kind == "SIREN":
        # fold the last activation like the kernel's autograd contract: dout is dL/d(out)
        grads_ref, dzs = O.siren_backward(sd, x, tr, dout, depth, net.get("last_tanh", False))
    else:
        masks = [eng.read_image("d", l, bs)[:bs].cpu() for l in range(depth - 1)]
        grads_ref, dzs = O.ffn_backward(sd, x, tr, dout, depth, masks=masks)
    print("NOTE external-dout path treats dout as dL/dz_last (last activation not applied)")
    g = eng.backward(dzs[depth - 1].cuda())
    torch.cuda.synchronize()
    sc = eng.scalars(bs)
    print("scalars", sc[:8].tolist())
    S = float(sc[1])
    for l in range(depth - 1):
        print(f"dZ{l} rel err", rel(eng.read_image("dz", l, bs)[:bs] / S, dzs[l]))
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        gv = g[off:off + rows * cols].view(grads_ref[k].shape)
        print(f"grad {k} rel err", rel(gv, grads_ref[k]), "norm", float(grads_ref[k].norm()))

    stage("fused train steps vs golden")
    eng2 = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
    eng2.load_tensors(list(sd.values()))
    eng2.set_encoder(encB)
    gold = G.load_golden(name)
    losses = []
    for step in range(G.N_ADAM_STEPS):
        eng2.train_step(loss_kind, cd, gd, bs, loss_opts=opts)
        torch.cuda.synchronize()
        losses.append(float(eng2.loss_out))
        if step == 0:
            print("step0 scalars", eng2.scalars(bs)[:8].tolist())
    print("losses engine", losses)
    print("losses golden", gold["losses"])
    for (off, rows, cols, layer, is_bias), k in zip(plan.tensors, sd.keys()):
        v = eng2.params[off:off + rows * cols]
        dg = G.tensor_digest(v.cpu())
        print(f"final {k}: l2 {dg['l2']:.8e} vs golden {gold['final'][k]['l2']:.8e}; head {dg['head'][:3]} vs {gold['final'][k]['head'][:3]}")

    stage("timing (bs 10000, eager launches)")
    bs2 = 10000
    plan2 = inr.Plan(model_kind, net, enc_cfg)
    eng3 = inr.ChainEngine(plan2, max_batch=bs2, lr=G.LR)
    eng3.load_tensors(list(sd.values()))
    eng3.set_encoder(encB)
    c2 = (torch.rand(bs2, 3, device="cuda") * 2 - 1)
    g2 = torch.rand(bs2, 2, device="cuda")
    for _ in range(5):
        eng3.train_step(loss_kind, c2, g2, bs2, loss_opts=opts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng3.train_step(loss_kind, c2, g2, bs2, loss_opts=opts)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"train_step bs={bs2}: {ms*1000:.1f} us/step -> {bs2/ms*1000:.3e} coords/s; loss {float(eng3.loss_out):.5f}")
    print("per-kernel ms:", eng3.profile_step(loss_kind, c2, g2, bs2, loss_opts=opts, reps=20))
    for bsx in (25000, 100000, 300000):
        engx = inr.ChainEngine(plan2, max_batch=bsx, lr=G.LR)
        engx.load_tensors(list(sd.values())); engx.set_encoder(encB)
        cx = (torch.rand(bsx, 3, device="cuda") * 2 - 1); gx = torch.rand(bsx, 2, device="cuda")
        print(f"bs={bsx} per-kernel ms:", engx.profile_step(loss_kind, cx, gx, bsx, loss_opts=opts, reps=10))
        del engx


if __name__ == "__main__":
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
