"""GPU parity tests of the WIRE2D path (reference src/models/wire2d.py) through the C ABI.  Same method as
tests/test_gpu_wire.py: the network is chaotic in fp32, so every stage is judged TEACHER-FORCED (recomputed in fp64
from the engine's own inputs to that stage, tolerance 1e-3 relative L2); end-to-end quantities are compared with the
fp64 run of the reference at the level the fp32 reference itself achieves."""
import pytest
import torch

from oracle import golden_util as G
from oracle import inr_oracle as O
from oracle.cases import case_setup, loss_and_grad

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a, b = a.cpu(), b.cpu()
    a = torch.view_as_real(a.to(torch.complex128)) if a.is_complex() else a.double()
    b = torch.view_as_real(b.to(torch.complex128)) if b.is_complex() else b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def to64(sd):
    return {k: (v.to(torch.complex128) if v.is_complex() else v.double()) for k, v in sd.items()}


@pytest.fixture(scope="module")
def inr():
    import mri_implicit_neural_representations_b200 as m
    return m


def _engine(inr):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("wire2d_l2")
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    return plan, eng, net, loss_kind, opts, sd, coords, gt


def _img(eng, kind, layer, bs, P, groups):
    """Decode an image family of `groups` x P features to a list of [bs, P] fp32 matrices (H: hi + lo parts summed)."""
    lay = eng.plan.workspace_layout(bs)
    T, F = lay["n_tiles"], groups * P
    nbytes = T * 128 * F * 2

    def dec(off):
        img = eng.workspace[off:off + nbytes].view(torch.float16).view(T, F // 8, 128, 8)
        return img.permute(0, 2, 1, 3).reshape(T * 128, F).float()

    if kind == "h":
        m = dec(lay["h"][layer]) + dec(lay["h"][layer] + nbytes)       # H_lo follows H_hi
    else:
        m = dec(lay["d" if kind == "z" else "dz"][layer])
    return [m[:bs, g * P:(g + 1) * P].cpu().double() for g in range(groups)]


def _act(lin, orth, w, s):
    la, lb = (lin.real, lin.imag) if lin.is_complex() else (lin, torch.zeros_like(lin))
    sa, sb = (orth.real, orth.imag) if orth.is_complex() else (orth, torch.zeros_like(orth))
    mag = torch.exp(-w * lb - s * s * (la * la + lb * lb + sa * sa + sb * sb))
    return torch.complex(mag * torch.cos(w * la), mag * torch.sin(w * la))


def test_wire2d_forward_teacher_forced_per_layer(inr):
    plan, eng, net, loss_kind, opts, sd, coords, gt = _engine(inr)
    depth, bs, C = net["network_depth"], coords.shape[0], net["network_width"]
    P = (C + 63) // 64 * 64
    sd64 = to64(sd)
    out32 = O.wire2d_forward(sd, coords, depth)
    tr64 = []
    out64 = O.wire2d_forward(sd64, coords.double(), depth, trace=tr64)
    out = eng.forward(coords.cuda(), train=True)
    hr, hi = _img(eng, "h", 1, bs, P, 2)
    assert rel(torch.complex(hr, hi)[:, :C], tr64[0][2]) <= 1e-5          # first layer (CUDA cores, fp32)
    for l in range(1, depth + 1):
        hr, hi = _img(eng, "h", l, bs, P, 2)
        hin = torch.complex(hr, hi)[:, :C]
        lin = hin @ sd64[f"net.{l}.linear.weight"].t() + sd64[f"net.{l}.linear.bias"]
        orth = hin @ sd64[f"net.{l}.scale_orth.weight"].t() + sd64[f"net.{l}.scale_orth.bias"]
        y = _act(lin, orth, float(sd[f"net.{l}.omega_0"]), float(sd[f"net.{l}.scale_0"]))
        yr, yi = _img(eng, "h", l + 1, bs, P, 2)
        assert rel(torch.complex(yr, yi)[:, :C], y) <= 5e-5, f"layer {l}"               # 3-pass split GEMM
        za, zb, zc, zd = _img(eng, "z", l, bs, P, 4)
        assert rel(torch.complex(za, zb)[:, :C], lin) <= TOL and rel(torch.complex(zc, zd)[:, :C], orth) <= TOL, l
    hr, hi = _img(eng, "h", depth + 1, bs, P, 2)
    o_tf = (torch.complex(hr, hi)[:, :C] @ sd64[f"net.{depth + 1}.weight"].t() + sd64[f"net.{depth + 1}.bias"]).real
    assert rel(out, o_tf) <= 1e-5
    assert rel(out, out64) <= 4 * rel(out32, out64) + 1e-4


def test_wire2d_backward_teacher_forced_and_gradients(inr):
    plan, eng, net, loss_kind, opts, sd, coords, gt = _engine(inr)
    depth, bs, C = net["network_depth"], coords.shape[0], net["network_width"]
    P = (C + 63) // 64 * 64
    L = depth + 1
    sd64 = to64(sd)
    out_dev = torch.zeros(bs, 2, device="cuda")
    eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts, out=out_dev)
    torch.cuda.synchronize()
    scal = eng.scalars(bs)
    Sl = scal[16:16 + depth + 1].tolist()
    val, dout = loss_and_grad(loss_kind, opts, out_dev.cpu().double(), gt.double(), coords.double())
    assert abs(float(eng.loss_out) - float(val)) <= TOL * abs(float(val))
    h = {}
    for l in range(1, L + 1):
        hr, hi = _img(eng, "h", l, bs, P, 2)
        h[l] = torch.complex(hr, hi)[:, :C]
    z = {l: [m[:, :C] for m in _img(eng, "z", l, bs, P, 4)] for l in range(0, depth + 1)}
    dz = {l: [m[:, :C] / Sl[l] for m in _img(eng, "dz", l, bs, P, 4)] for l in range(0, depth + 1)}

    def act_grad(dh, y, zl, l):
        w_, s2 = float(sd[f"net.{l}.omega_0"]), float(sd[f"net.{l}.scale_0"]) ** 2
        pq = dh.conj() * y
        za, zb, zc, zd = zl
        da = -2 * s2 * za * pq.real - w_ * pq.imag
        dc = -2 * s2 * zc * pq.real
        if l == 0:
            return [da, torch.zeros_like(da), dc, torch.zeros_like(da)]
        return [da, -(w_ + 2 * s2 * zb) * pq.real, dc, -2 * s2 * zd * pq.real]

    def cat(parts):
        return torch.cat(parts, dim=1)

    dh = torch.complex(dout, torch.zeros_like(dout)) @ sd64[f"net.{L}.weight"].conj()
    assert rel(cat(dz[depth]), cat(act_grad(dh, h[L], z[depth], depth))) <= TOL
    for l in range(depth, 0, -1):
        dlin, dorth = torch.complex(dz[l][0], dz[l][1]), torch.complex(dz[l][2], dz[l][3])
        dh = dlin @ sd64[f"net.{l}.linear.weight"].conj() + dorth @ sd64[f"net.{l}.scale_orth.weight"].conj()
        assert rel(cat(dz[l - 1]), cat(act_grad(dh, h[l], z[l - 1], l - 1))) <= TOL, f"dZ{l-1}"
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    for l in range(1, depth + 1):
        dlin, dorth = torch.complex(dz[l][0], dz[l][1]), torch.complex(dz[l][2], dz[l][3])
        assert rel(gv[f"net.{l}.linear.weight"], dlin.t() @ h[l].conj()) <= TOL, l
        assert rel(gv[f"net.{l}.linear.bias"], dlin.sum(0)) <= TOL, l
        assert rel(gv[f"net.{l}.scale_orth.weight"], dorth.t() @ h[l].conj()) <= TOL, l
        assert rel(gv[f"net.{l}.scale_orth.bias"], dorth.sum(0)) <= TOL, l
    x64 = coords.double()
    assert rel(gv["net.0.linear.weight"], dz[0][0].t() @ x64) <= TOL
    assert rel(gv["net.0.linear.bias"], dz[0][0].sum(0)) <= TOL
    assert rel(gv["net.0.scale_orth.weight"], dz[0][2].t() @ x64) <= TOL
    assert rel(gv["net.0.scale_orth.bias"], dz[0][2].sum(0)) <= TOL
    dzl = torch.complex(dout, torch.zeros_like(dout))
    assert rel(gv[f"net.{L}.weight"], dzl.t() @ h[L].conj()) <= TOL
    assert rel(gv[f"net.{L}.bias"].real, dout.sum(0)) <= TOL
    # end to end against fp64 autograd: inside the drift band of the fp32 reference itself
    Pm = {k: v.clone().requires_grad_(not (k.endswith("omega_0") or k.endswith("scale_0"))) for k, v in sd64.items()}
    o64 = O.wire2d_forward(Pm, x64, depth)
    _, d64 = loss_and_grad(loss_kind, opts, o64.detach(), gt.double(), x64)
    live = [k for k in Pm if Pm[k].requires_grad]
    ref = dict(zip(live, torch.autograd.grad(o64, [Pm[k] for k in live], grad_outputs=d64)))
    for k in live:
        assert rel(gv[k], ref[k]) <= 5e-2, (k, rel(gv[k], ref[k]))


def test_wire2d_fused_steps_module_and_frozen_parameters(inr):
    plan, eng, net, loss_kind, opts, sd, coords, gt = _engine(inr)
    bs = coords.shape[0]
    gold = G.load_golden("wire2d_l2")
    before = {k: v.clone() for k, v in zip(sd.keys(), eng.param_views())}
    losses = []
    for step in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
        losses.append(float(eng.loss_out))
    assert abs(losses[0] - gold["fp64"]["losses"][0]) <= 5e-4 * gold["fp64"]["losses"][0]
    lo = min(gold["losses"][-1], gold["fp64"]["losses"][-1])
    hi = max(gold["losses"][-1], gold["fp64"]["losses"][-1])
    assert 0.9 * lo <= losses[-1] <= 1.1 * hi
    after = dict(zip(sd.keys(), eng.param_views()))
    for k in sd:
        if k.endswith("omega_0") or k.endswith("scale_0"):
            assert torch.equal(after[k].cpu(), before[k].cpu()), k
        else:
            assert not torch.equal(after[k].cpu(), before[k].cpu()), k
    # drop-in module: reference key order, forward equals the engine's
    from mri_implicit_neural_representations_b200.modules import WIRE2D
    torch.manual_seed(3)
    m = WIRE2D(dict(net)).to("cuda")
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict({k: v for k, v in sd.items()})
    with torch.no_grad():
        o = m(coords.cuda())
    assert rel(o, O.wire2d_forward(to64(sd), coords.double(), net["network_depth"])) <= 5e-2


# ------------------------------------------------------------------------------------------------ complex tanh tail
def _act_grad(sd, dh, y, zl, l):
    w_, s2 = float(sd[f"net.{l}.omega_0"]), float(sd[f"net.{l}.scale_0"]) ** 2
    pq = dh.conj() * y
    za, zb, zc, zd = zl
    da = -2 * s2 * za * pq.real - w_ * pq.imag
    dc = -2 * s2 * zc * pq.real
    if l == 0:
        return [da, torch.zeros_like(da), dc, torch.zeros_like(da)]
    return [da, -(w_ + 2 * s2 * zb) * pq.real, dc, -2 * s2 * zd * pq.real]


def _engine_tanh(inr):
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, coords, gt, mask = case_setup("wire2d_tanh")
    plan = inr.Plan(model_kind, net, enc_cfg)
    eng = inr.ChainEngine(plan, max_batch=coords.shape[0], lr=G.LR)
    eng.load_tensors(list(sd.values()))
    return plan, eng, net, loss_kind, opts, sd, coords, gt


def test_wire2d_tanh_tail_forward_backward_teacher_forced(inr):
    """reference wire2d.py:106-107,115-116: torch.nn.Tanh on the complex final linear, then `.real`.  The final stage is judged
    teacher-forced on the engine's own last-layer activations (fp64): output, the gradient handed to the last hidden layer,
    and the COMPLEX gradients of the final weight / bias (the imaginary parts exist only because of the tail)."""
    plan, eng, net, loss_kind, opts, sd, coords, gt = _engine_tanh(inr)
    depth, bs, C = net["network_depth"], coords.shape[0], net["network_width"]
    P = (C + 63) // 64 * 64
    L = depth + 1
    sd64 = to64(sd)
    out_dev = torch.zeros(bs, 2, device="cuda")
    eng.grad_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts, out=out_dev)
    torch.cuda.synchronize()
    Sl = eng.scalars(bs)[16:16 + depth + 1].tolist()
    hr, hi = _img(eng, "h", L, bs, P, 2)
    hL = torch.complex(hr, hi)[:, :C]
    z = (hL @ sd64[f"net.{L}.weight"].t() + sd64[f"net.{L}.bias"]).requires_grad_(True)
    o_tf = torch.tanh(z).real
    assert rel(out_dev, o_tf.detach()) <= 1e-5
    val, dout = loss_and_grad(loss_kind, opts, out_dev.cpu().double(), gt.double(), coords.double())
    assert abs(float(eng.loss_out) - float(val)) <= TOL * abs(float(val))
    gz = torch.autograd.grad(o_tf, z, grad_outputs=dout)[0]          # dL/dx + j dL/dy
    assert float(gz.imag.abs().max()) > 1e-3 * float(gz.real.abs().max())      # the tail does feed the imaginary part
    dh = gz @ sd64[f"net.{L}.weight"].conj()
    zl = [m[:, :C] for m in _img(eng, "z", depth, bs, P, 4)]
    dzl = [m[:, :C] / Sl[depth] for m in _img(eng, "dz", depth, bs, P, 4)]
    assert rel(torch.cat(dzl, 1), torch.cat(_act_grad(sd, dh, hL, zl, depth), 1)) <= TOL
    gv = dict(zip(sd.keys(), eng._views(eng.grads)))
    assert rel(gv[f"net.{L}.weight"], gz.t() @ hL.conj()) <= TOL
    assert rel(gv[f"net.{L}.bias"], gz.sum(0)) <= TOL
    # end to end against fp64 autograd on the oracle: inside the drift band of the fp32 reference itself
    x64 = coords.double()
    Pm = {k: v.clone().requires_grad_(not (k.endswith("omega_0") or k.endswith("scale_0"))) for k, v in sd64.items()}
    o64 = O.model_forward("WIRE2D", Pm, x64, net)
    o32 = O.model_forward("WIRE2D", sd, coords, net)
    assert rel(out_dev, o64.detach()) <= 4 * rel(o32, o64.detach()) + 1e-4
    _, d64 = loss_and_grad(loss_kind, opts, o64.detach(), gt.double(), x64)
    live = [k for k in Pm if Pm[k].requires_grad]
    ref = dict(zip(live, torch.autograd.grad(o64, [Pm[k] for k in live], grad_outputs=d64)))
    for k in live:
        assert rel(gv[k], ref[k]) <= 5e-2, (k, rel(gv[k], ref[k]))


def test_wire2d_tanh_tail_fused_steps_and_module(inr):
    plan, eng, net, loss_kind, opts, sd, coords, gt = _engine_tanh(inr)
    bs = coords.shape[0]
    gold = G.load_golden("wire2d_tanh")
    losses = []
    for step in range(G.N_ADAM_STEPS):
        eng.train_step(loss_kind, coords.cuda(), gt.cuda(), bs, loss_opts=opts)
        losses.append(float(eng.loss_out))
    assert abs(losses[0] - gold["fp64"]["losses"][0]) <= 5e-4 * gold["fp64"]["losses"][0]
    lo = min(gold["losses"][-1], gold["fp64"]["losses"][-1])
    hi = max(gold["losses"][-1], gold["fp64"]["losses"][-1])
    assert 0.9 * lo <= losses[-1] <= 1.1 * hi
    # drop-in module (autograd face): output and parameter .grad against torch autograd on the fp64 oracle
    from mri_implicit_neural_representations_b200.modules import WIRE2D
    torch.manual_seed(3)
    m = WIRE2D(dict(net)).to("cuda")
    m.load_state_dict({k: v for k, v in sd.items()})
    o = m(coords.cuda())
    sd64 = to64(sd)
    Pm = {k: v.clone().requires_grad_(not (k.endswith("omega_0") or k.endswith("scale_0"))) for k, v in sd64.items()}
    o64 = O.model_forward("WIRE2D", Pm, coords.double(), net)
    assert rel(o.detach(), o64.detach()) <= 5e-2
    w = torch.linspace(0.5, 1.5, bs * 2).view(bs, 2)
    (o * w.cuda()).sum().backward()
    (o64 * w.double()).sum().backward()
    for k, p in m.state_dict(keep_vars=True).items():
        if Pm[k].requires_grad:
            assert rel(p.grad, Pm[k].grad) <= 5e-2, k
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k


def test_wire2d_full_size_batch_properties(inr):
    """Bench workload size (WIRE2D d8 w256, bs 25 000 = 196 row tiles through the chained forward / dgrad launches; reference
    wire2d.py:4-118).  Rows sampled from all over the batch match the fp64 oracle as well as the fp32 oracle does; a row
    permutation of the batch leaves every row's output bit-identical and loss / gradients unchanged up to the order of the
    fixed-order reductions; the same step twice gives the same bits."""
    model_kind, net, enc_cfg, loss_kind, opts, sd, encB, _, _, _ = case_setup("wire2d_l2")
    bs = 25000
    g = torch.Generator().manual_seed(5)
    c = torch.rand(bs, 3, generator=g) * 2 - 1
    y = torch.randn(bs, 2, generator=g) * 0.05

    def run(c_, y_):
        eng = inr.ChainEngine(inr.Plan(model_kind, net, enc_cfg), max_batch=bs, lr=G.LR)
        eng.load_tensors(list(sd.values()))
        out = torch.zeros(bs, 2, device="cuda")
        cd, yd = c_.cuda(), y_.cuda()
        for _ in range(2):                                       # second pass: lagged per-layer scales calibrated
            eng.grad_step(loss_kind, cd, yd, bs, loss_opts=opts, out=out)
        torch.cuda.synchronize()
        return out.cpu(), float(eng.loss_out), eng.grads.clone()

    out, loss, grads = run(c, y)
    rows = torch.cat([torch.arange(0, 200), torch.randint(200, bs - 200, (300,), generator=g), torch.arange(bs - 200, bs)])
    o32 = O.model_forward(model_kind, sd, c[rows], net)
    o64 = O.model_forward(model_kind, to64(sd), c[rows].double(), net)
    assert rel(out[rows], o64) <= 4 * rel(o32, o64) + 1e-4
    out2, loss2, grads2 = run(c, y)
    assert torch.equal(out2, out) and loss2 == loss and torch.equal(grads2, grads)
    perm = torch.randperm(bs, generator=g)
    out_p, loss_p, grads_p = run(c[perm], y[perm])
    assert torch.equal(out_p, out[perm])
    assert abs(loss_p - loss) <= 1e-5 * abs(loss), (loss_p, loss)
    assert float((grads_p - grads).norm()) <= 1e-4 * float(grads.norm())
    assert float(grads.norm()) > 0 and torch.isfinite(grads).all()
