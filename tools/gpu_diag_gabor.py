"""Gabor gradient diagnosis: per-tensor relative errors and the q-reductions (Qx, s, u) against fp64 torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mri_implicit_neural_representations_b200 as inr
from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup, loss_and_grad


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


cond = len(sys.argv) < 2 or sys.argv[1] != "ref"
model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("gabor_tanh")
L_ = net["network_depth"]
if cond:
    for i in range(L_ + 1):
        sd[f"filters.{i}.gamma"] = sd[f"filters.{i}.gamma"] * 0.01 + 1e-3
        sd[f"filters.{i}.mu"] = sd[f"filters.{i}.mu"] * 0.5
plan = inr.Plan("Gabor", net, enc_cfg)
bs = coords.shape[0]
eng = inr.ChainEngine(plan, max_batch=bs, lr=G.LR)
eng.load_tensors(list(sd.values()))
eng.set_encoder(encB)
x = O.encode(coords, encB, "gauss")
P = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
x64 = x.double()
fs = []
z = O._filter(P, 0, x64, True); fs.append(z)
z.retain_grad()
zz = z
for i in range(1, L_ + 1):
    f = O._filter(P, i, x64, True); f.retain_grad(); fs.append(f)
    zz = f * (zz @ P[f"linear.{i-1}.weight"].t() + P[f"linear.{i-1}.bias"])
o = zz @ P["output_linear.weight"].t() + P["output_linear.bias"]
val, dout = loss_and_grad(loss_kind, opts, o.detach().float(), gt, coords)
o.backward(dout.double())
for rep in range(2):
    eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
print("loss", float(eng.loss_out), float(val))
gv = dict(zip(sd.keys(), eng._views(eng.grads)))
for k in sd:
    print(f"{k:28s} rel {rel(gv[k], P[k].grad):.3e}   |ref| {float(P[k].grad.norm()):.3e}")
lay = plan.workspace_layout(bs)
sc = eng.scalars(bs)
W, IN = net["network_width"], net["network_input_size"]
gp = eng.workspace[lay["gpart"]:lay["gpart"] + lay["n_split"] * lay["gstride"] * 4].view(torch.float32).view(lay["n_split"], lay["gstride"])
for i in (0, 4, L_):
    S = float(sc[16 + i])
    q_ref = (fs[i].grad * fs[i]).detach()
    q = eng.read_mfn_image("q", i, bs)[:bs].cpu().double() / S
    t = [tt for tt in plan.tensors]
    names = list(sd.keys())
    mu_off = t[names.index(f"filters.{i}.mu")][0]
    Qx = gp[:, mu_off:mu_off + W * IN].sum(0).view(W, IN).cpu().double() / S
    aux = gp[:, lay["aux0"] + i * W * 16: lay["aux0"] + (i + 1) * W * 16].sum(0).view(W, 16).cpu().double() / S
    xn = (x64 ** 2).sum(-1)
    print(f"stage {i}: S={S:g} q rel {rel(q, q_ref):.3e}  Qx rel {rel(Qx, q_ref.t() @ x64):.3e}  s rel {rel(aux[:, 0], q_ref.sum(0)):.3e}"
          f"  u rel {rel(aux[:, 1], (q_ref * xn[:, None]).sum(0)):.3e}   |Qx| {float((q_ref.t() @ x64).norm()):.3e} |s mu| {float((q_ref.sum(0)[:, None] * P[f'filters.{i}.mu']).norm()):.3e}")
