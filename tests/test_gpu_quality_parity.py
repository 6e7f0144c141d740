"""End-to-end quality parity (north_star: final PSNR / SSIM within 0.1 dB / 0.002 of the reference after a fixed step
count) of the product entry points against the ORACLE's own training loop -- forward, closed-form loss gradient, torch
autograd through the oracle model, explicit Adam, per-epoch learning-rate decay -- from the same seed on the same small
synthetic slice, short horizon, for the BASELINE configurations the round-1 suite did not cover:

  config 2  WIRE complex Gabor, k-space, HDR loss, undersampling grid-2*1 (src/train.py through FusedTrainer)
  config 3  FFN and Gabor MFN, k-space, tanh loss, per-coil batches + TV (fused TV kernel) -- against the oracle, not
            against the engine's own autograd face
  config 4  MultiscaleBoundedFourier, LSL + 0.1 ConsistencyLoss (src/train_kspace_multiscale.py through the fused step)

The oracle is pinned against the unmodified reference modules / losses (tests/test_oracle_vs_reference.py,
tests/test_losses_vs_reference.py)."""
import os
import sys
import warnings
from collections import OrderedDict

import pytest
import torch

from oracle import inr_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "src")
ENC = {"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3}


@pytest.fixture(scope="module")
def src_path():
    sys.path.insert(0, SRC)
    yield
    sys.path.remove(SRC)


def _cfg(model, net, enc, loss, bs, epochs, lr=5e-4, **kw):
    c = {"model": model, "net": dict(net), "encoder": dict(enc), "loss": loss, "optimizer": "Adam", "lr": lr, "beta1": 0.9,
         "beta2": 0.999, "weight_decay": 0.0, "max_epoch": epochs, "batch_size": bs, "log_iter": 1000, "val_epoch": epochs,
         "image_save_epoch": 10 ** 6, "transform": False, "data": "knee", "regularization": {"type": "none"}, "per_coil": False,
         "use_tv": False, "undersampling": None, "loss_opts": {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}}
    c.update(kw)
    return c


def _quality(ds, flat):
    C, H, W, _ = ds.img_shape
    gt = O.rss(O.complex_abs(O.ifft2c(ds.image.reshape(C, H, W, 2))), 0)
    rec = O.rss(O.complex_abs(O.ifft2c(flat.reshape(C, H, W, 2))), 0)
    return float(O.psnr(gt, rec)), float(O.ssim(gt.numpy(), rec.numpy()))


def _oracle_fit(kind, net, enc_cfg, ds, bs, epochs, lr0, loss, opts, mask=None, tv=None, forward=None):
    """The reference loop body (src/train.py:158-192) on the oracle: grid-order batches, row mask, optional per-coil TV,
    Adam with the per-epoch LambdaLR decay (:153,:251).  Returns the final state_dict and encoder matrix."""
    encB = O.encoder_init(enc_cfg)
    sd = O.MODEL_INIT[kind](dict(net))
    frozen = {k for k in sd if k.endswith("omega_0") or k.endswith("scale_0")}
    P = OrderedDict((k, v.clone()) for k, v in sd.items())
    state = {k: (torch.zeros_like(torch.view_as_real(v) if v.is_complex() else v),
                 torch.zeros_like(torch.view_as_real(v) if v.is_complex() else v)) for k, v in P.items() if k not in frozen}
    t = 0
    n = len(ds)
    for e in range(epochs):
        lr = O.lr_at_epoch(lr0, e, epochs)
        for i in range(0, n, bs):
            t += 1
            c, y = ds.coords[i:i + bs], ds.image[i:i + bs]
            leafs = OrderedDict((k, v.clone().requires_grad_(k not in frozen)) for k, v in P.items())
            out = O.model_forward(kind, leafs, O.encode(c, encB, enc_cfg["embedding"]), net)
            mb = None if mask is None else mask[i:i + bs]
            o_sel, y_sel = (out, y) if mb is None else (out[mb], y[mb])
            if loss == "HDR":
                _, g, _ = O.loss_hdr(o_sel.detach(), y_sel, c, float(opts["hdr_ff_sigma"]), float(opts["hdr_eps"]), float(opts["hdr_ff_factor"]))
            else:
                _, g = O.LOSS_TRAIN[loss](o_sel.detach(), y_sel)
            g_full = torch.zeros_like(out)
            if mb is None:
                g_full = g
            else:
                g_full[mb] = g
            if tv is not None and mb is not None:                   # src/train.py:172-174: TV on ALL rows of the coil
                _, g_tv = O.loss_tv(out.detach(), tv[0], tv[1])
                g_full = g_full + g_tv
            live = [p for k, p in leafs.items() if k not in frozen]
            grads = torch.autograd.grad(out, live, grad_outputs=g_full)
            for (k, p), gr in zip([(k, p) for k, p in P.items() if k not in frozen], grads):
                pr = torch.view_as_real(p) if p.is_complex() else p
                gr = torch.view_as_real(gr.contiguous()) if gr.is_complex() else gr
                O.adam_step(pr, gr, state[k][0], state[k][1], t, lr)
    return P, encB


def _loaders(bs, shape, **kw):
    from data.slices import get_data_loader
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return get_data_loader("knee", "data", "train", bs, transform=False, normalization="max", shape=shape, **kw)


WIRE_NET = {"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256, "first_omega_0": 30,
            "hidden_omega_0": 30, "scale": 15}
WIRE_ENC = {"embedding": "none", "scale": 4, "embedding_size": 256, "coordinates_size": 3}


def _wire_runs(tmp_path, epochs, oracle_threads, bs=2000):
    """Engine run through src/train.py + one oracle run per thread count (same seed, same batches)."""
    import train
    shape = (4, 48, 40)                                             # 7680 rows; bs 2000 -> 4 batches / epoch, the last one 1680
    ds, tl, vl = _loaders(bs, shape, undersampling="grid-2*1")
    cfg = _cfg("WIRE", WIRE_NET, WIRE_ENC, "HDR", bs, epochs, undersampling="grid-2*1")
    torch.manual_seed(91)
    hist = train.training_script(cfg, ds, tl, vl, 0, 0, output_path=str(tmp_path), verbose=False)
    _, psnr_e, ssim_e = hist[-1]
    tds = tl.ds                    # the training view of the slice: undersampled targets + row mask (validation uses `ds`)
    runs = []
    n0 = torch.get_num_threads()
    try:
        for nt in oracle_threads:
            torch.set_num_threads(nt)
            torch.manual_seed(91)
            P, _ = _oracle_fit("WIRE", WIRE_NET, WIRE_ENC, tds, bs, epochs, 5e-4, "HDR", cfg["loss_opts"], mask=tds.coords_mask[:, 0])
            with torch.no_grad():
                flat = O.model_forward("WIRE", P, ds.coords, WIRE_NET)
            runs.append(_quality(ds, flat))
    finally:
        torch.set_num_threads(n0)
    return (psnr_e, ssim_e), runs


def test_config2_wire_hdr_masked_first_step_vs_fp64_oracle(src_path, tmp_path):
    """Short horizon for WIRE + HDR, judged where it can be judged.  Measured on this slice (real grid coordinates, depth 4):
    the network is ill-conditioned at its initialisation -- the fp32 ORACLE's forward is 1.3e-3 from the fp64 oracle (each
    wavelet multiplies a pre-activation error by ~omega + 2 sigma^2 |z|), the engine's 3.2e-3 -- and HDR's log-ratio turns
    that into ~1 % of the loss value; after two sign-like Adam steps the engine sits 0.23 dB from one oracle run, and oracle
    runs that differ only in CPU thread count spread over 2.88 .. 2.92 dB / SSIM 0.091 .. 0.106 after four steps.  PSNR / SSIM
    at ~2.8 dB are noise there; the long-horizon band test below carries that claim.  Here:
      * forward of the first batch: engine vs fp64 oracle no further than 4x the fp32 oracle's own distance (+1e-4);
      * the fused HDR + mask loss equals the oracle's HDR loss evaluated on the ENGINE's output (teacher-forced, 1e-5);
      * and lies within 3 % of the fp64 loss."""
    import mri_implicit_neural_representations_b200 as inr
    bs = 2000
    ds, tl, vl = _loaders(bs, (4, 48, 40), undersampling="grid-2*1")
    tds = tl.ds
    opts = {"hdr_eps": 1e-2, "hdr_ff_sigma": 1.0, "hdr_ff_factor": 0.5}
    torch.manual_seed(91)
    sd = O.wire_init(dict(WIRE_NET))
    c, y, mb = tds.coords[:bs], tds.image[:bs], tds.coords_mask[:bs, 0]
    eng = inr.ChainEngine(inr.Plan("WIRE", WIRE_NET, WIRE_ENC), max_batch=bs, lr=5e-4)
    eng.load_tensors(list(sd.values()))
    out = torch.empty(bs, 2, device="cuda")
    eng.grad_step("HDR", c.cuda(), y.cuda(), bs, mask=mb.to(torch.uint8).cuda(), loss_opts=opts, out=out)
    out, loss_e = out.cpu(), float(eng.loss_out)
    o32 = O.model_forward("WIRE", sd, c, WIRE_NET)
    sd64 = OrderedDict((k, v.to(torch.complex128 if v.is_complex() else torch.float64)) for k, v in sd.items())
    o64 = O.model_forward("WIRE", sd64, c.double(), WIRE_NET)
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert rel(out, o64) <= 4 * rel(o32, o64) + 1e-4, (rel(out, o64), rel(o32, o64))
    l_tf = float(O.loss_hdr(out[mb], y[mb], c, 1.0, 1e-2, 0.5)[0])
    l64 = float(O.loss_hdr(o64[mb], y.double()[mb], c.double(), 1.0, 1e-2, 0.5)[0])
    assert abs(loss_e - l_tf) <= 1e-5 * l_tf, (loss_e, l_tf)
    assert abs(loss_e - l64) <= 3e-2 * l64, (loss_e, l64)


def test_config2_wire_hdr_masked_long_horizon_band(src_path, tmp_path):
    """SURVEY 8c-5 long-horizon protocol: WIRE's fp32 trajectory is chaotic -- the reference arithmetic itself decorrelates
    when only the summation order changes -- so after 24 steps the engine must land inside the min-max band of three
    oracle runs that differ ONLY in their CPU thread count (i.e. in summation order), widened by the 0.1 dB / 0.002
    tolerance.  The band is part of the assertion message."""
    (psnr_e, ssim_e), runs = _wire_runs(tmp_path, 6, [1, 3, 8])
    p_lo, p_hi = min(r[0] for r in runs), max(r[0] for r in runs)
    s_lo, s_hi = min(r[1] for r in runs), max(r[1] for r in runs)
    assert p_lo - 0.1 <= psnr_e <= p_hi + 0.1, (psnr_e, runs)
    assert s_lo - 0.002 <= ssim_e <= s_hi + 0.002, (ssim_e, runs)


@pytest.mark.parametrize("model_name", ["FFN", "Gabor"])
def test_config3_tanh_tv_per_coil_quality_parity_vs_oracle(src_path, tmp_path, model_name):
    import train
    net = {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256}
    if model_name == "Gabor":
        net.update(network_depth=2)
    shape, epochs = (3, 24, 32), 3
    bs = shape[1] * shape[2]
    ds, tl, vl = _loaders(bs, shape, undersampling="grid-2*1", per_coil=True)
    cfg = _cfg(model_name, net, ENC, "tanh", bs, epochs, per_coil=True, use_tv=True, undersampling="grid-2*1")
    torch.manual_seed(31)
    hist = train.training_script(cfg, ds, tl, vl, 0, 0, output_path=str(tmp_path), verbose=False)
    _, psnr_e, ssim_e = hist[-1]
    torch.manual_seed(31)
    tds = tl.ds
    P, encB = _oracle_fit(model_name, net, ENC, tds, bs, epochs, 5e-4, "tanh", None, mask=tds.coords_mask[:, 0], tv=(shape[1], shape[2]))
    with torch.no_grad():
        flat = O.model_forward(model_name, P, O.encode(ds.coords, encB, "gauss"), net)
    psnr_o, ssim_o = _quality(ds, flat)
    assert abs(psnr_e - psnr_o) <= 0.1, (psnr_e, psnr_o)
    assert abs(ssim_e - ssim_o) <= 0.002, (ssim_e, ssim_o)


def test_config4_bounded_fourier_lsl_consistency_quality_parity(src_path, tmp_path):
    import train_kspace_multiscale as TM
    from clustering import partition_and_stats
    net = {"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 256}
    shape, bs, epochs = (2, 64, 64), 4096, 3
    ds, tl, vl = _loaders(bs, shape, use_dists="yes")
    cfg = _cfg("BoundedFourier", net, ENC, "LSL", bs, epochs, partition={"no_steps": 16, "no_models": 4},
               loss_opts={"hdr_ff_sigma": 1.0, "hdr_eps": 1e-2, "hdr_ff_factor": 0.0})
    torch.manual_seed(3)
    hist = TM.training_multiscale(cfg, ds, tl, vl, output_path=str(tmp_path), verbose=False)
    _, _, psnr_e, ssim_e = hist[-1]
    # oracle: same partition (clustering is pinned against the reference), same seed -> encoder B, then the model
    _, radii = partition_and_stats(dataset=ds, no_steps=16, no_parts=4, stat="max", show=False)
    pairs = [(float(a), float(b)) for a, b in TM.create_pairs(radii, 1)]
    bounds = [(float(a), float(b)) for a, b in TM.create_pairs(radii, 2)]
    torch.manual_seed(3)
    encB = O.encoder_init(ENC)
    sd = O.multiscale_init(dict(net), bounded=True)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(P.values()), lr=5e-4)
    dist = ds.dist_to_center
    for e in range(epochs):
        for g_ in opt.param_groups:
            g_["lr"] = O.lr_at_epoch(5e-4, e, epochs)
        for i in range(0, len(ds), bs):
            c, y, d = ds.coords[i:i + bs], ds.image[i:i + bs], dist[i:i + bs]
            outs = O.multiscale_forward(P, O.encode(c, encB, "gauss"), 8, d, bounds)
            _, douts = O.loss_consistency([o.detach() for o in outs], d, pairs, 0.1)
            for k, o in enumerate(outs):
                _, g = O.loss_logspace(o.detach(), y, 1e-2)
                douts[k] = douts[k] + g
            opt.zero_grad()
            torch.autograd.backward(outs, douts)
            opt.step()
    with torch.no_grad():
        flat = O.multiscale_forward(P, O.encode(ds.coords, encB, "gauss"), 8, dist, bounds)[-1]
    psnr_o, ssim_o = _quality(ds, flat)
    assert abs(psnr_e - psnr_o) <= 0.1, (psnr_e, psnr_o)
    assert abs(ssim_e - ssim_o) <= 0.002, (ssim_e, ssim_o)
