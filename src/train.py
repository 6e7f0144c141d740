"""Training script -- drop-in for the reference's src/train.py (same CLI, config keys, checkpoint layout).

The inner batch loop (reference :158-192) is replaced by the B200 engine:
  * fusable model + loss (+ regulariser): FusedTrainer -- encoder, model, loss, backward and Adam in one CUDA
    graph per batch, inputs resident on the device, loss read back only every log_iter steps;
  * anything else: the reference's own loop shape (model(x) -> loss -> backward -> optim.step) on the fused
    model's autograd face.
"""
import argparse
import os
import shutil
import sys
from datetime import datetime

import torch
from torch.optim.lr_scheduler import LambdaLR

_SRC = os.path.dirname(os.path.abspath(__file__))
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

from models.networks import FFN, SIREN, WIRE, Positional_Encoder            # noqa: E402
from models.mfn import FourierNet, GaborNet, KGaborNet                       # noqa: E402
from models.regularization import Regularization_L1, Regularization_L2       # noqa: E402
from metrics.losses import (CenterLoss, FocalFrequencyLoss, HDRLoss_FF, MSLELoss, TanhL2Loss, TLoss,    # noqa: E402
                            tv_loss)
from data.slices import get_data_loader                                      # noqa: E402
from log_handler.logger import INRLogger                                     # noqa: E402
from utils import get_config, set_default_configs                            # noqa: E402
from mri_implicit_neural_representations_b200 import metrics as M            # noqa: E402
from mri_implicit_neural_representations_b200.trainer import (FUSABLE_LOSSES, DataParallel, FusedAdam,   # noqa: E402
                                                              FusedTrainer)

opts = None      # the reference reads a module-global `opts` (src/train.py:35,45-48); kept for callers that set it


def get_device(net_name=None):
    if not torch.cuda.is_available():
        raise RuntimeError("this engine has no CPU fallback: a CUDA device (B200) is required")
    return torch.device("cuda", torch.cuda.current_device())


def init_distributed():
    """One process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment): binds the
    process to its GPU and joins the NCCL group.  Returns (rank, world); (0, 1) for a plain `python src/train.py` run.
    The reference has no multi-GPU path (its loop is single-device, src/train.py:27-252); see `run_fits` below for how
    the work is spread."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo",
                                device_id=torch.device("cuda", local) if torch.cuda.is_available() else None)
    return dist.get_rank(), dist.get_world_size()


def plan_fits(fits, rank, world, parallel="auto"):
    """How a list of (sample, slice) fits is spread over `world` ranks:
      * 'independent' (auto when there are at least as many fits as ranks): fit i runs on rank i % world, whole and on its
        own -- no data-path collective, exactly the reference's sequential loop (src/train.py:292-318) cut into `world` lanes;
      * 'dp' (auto otherwise, i.e. a single slice): every rank works on every fit, coordinate data-parallel.
    Returns (mode, fits of this rank)."""
    fits = list(fits)
    if world <= 1:
        return "single", fits
    if parallel == "auto":
        parallel = "independent" if len(fits) >= world else "dp"
    if parallel == "independent":
        return "independent", fits[rank::world]
    return "dp", fits


def build_model(config):
    name = config["model"]
    if name == "SIREN":
        return SIREN(config["net"])
    if name == "FFN":
        return FFN(config["net"])
    if name == "WIRE":
        return WIRE(config["net"])
    if name == "Fourier":
        return FourierNet(config["net"])
    if name == "Gabor":
        return GaborNet(config["net"])
    if name == "KGabor":
        return KGaborNet(config["net"])
    if name == "WIRE2D":                       # reference src/train.py:59-60, hp_model_training.py:55-56
        from models.wire2d import WIRE2D
        return WIRE2D(config["net"])
    raise NotImplementedError(name)            # reference :69-70


def build_loss(config):
    loss = config["loss"]
    if loss == "L2":
        return torch.nn.MSELoss()
    if loss == "MSLE":
        return MSLELoss()
    if loss == "L1":
        return torch.nn.L1Loss()
    if loss == "HDR":
        return HDRLoss_FF(config["loss_opts"])
    if loss == "tanh":
        return TanhL2Loss()
    if loss == "T":
        return TLoss()
    if loss == "LSL":
        return CenterLoss(config["loss_opts"])     # src/train.py maps LSL to CenterLoss (:87-88), not to LogSpaceLoss
    if loss == "FFL":
        return FocalFrequencyLoss()
    return None                                # reference falls through silently (:97-98)


def training_script(config, dataset, data_loader, val_loader, sample, slice_no, output_path=None, config_path=None,
                    verbose=True, dp=None):
    """Same positional signature as the reference.  Returns the list of (epoch, psnr, ssim) validations.
    dp: a trainer.DataParallel context -> this fit runs coordinate data-parallel over its ranks (rank 0 logs and saves)."""
    max_epoch = config["max_epoch"]
    in_image_space = config["transform"]
    device = get_device(config["model"])
    out_root = output_path or (opts.output_path if opts else ".")
    cfg_path = config_path or (opts.config if opts else None)
    tag = os.path.splitext(os.path.basename(cfg_path))[0] if cfg_path else "config"
    model_name = os.path.join(tag, "{}/img_sample{}_slice{}_{}_{}_{}_{}_{}_lr{:.2g}_encoder_{}".format(
        config["data"], sample, slice_no, config["model"], config["net"]["network_input_size"],
        config["net"]["network_width"], config["net"]["network_depth"], config["loss"], config["lr"],
        config["encoder"]["embedding"]))
    if config["encoder"]["embedding"] != "none":            # reference :40-41
        model_name += "_scale{}_size{}".format(config["encoder"]["scale"], config["encoder"]["embedding_size"])
    model_name += datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
    writes = dp is None or dp.rank == 0                      # data-parallel replicas are identical: one rank logs and saves
    train_writer = INRLogger(os.path.join(out_root, "logs", model_name)) if writes else None
    checkpoint_directory = os.path.join(out_root, "outputs", model_name, "checkpoints") if writes else None
    if writes:
        os.makedirs(checkpoint_directory, exist_ok=True)
        if cfg_path and os.path.exists(cfg_path):
            shutil.copy(cfg_path, os.path.join(out_root, "outputs", model_name, "config.yaml"))

    return fit(config, dataset, data_loader, val_loader, max_epoch, device, train_writer=train_writer,
               checkpoint_directory=checkpoint_directory, verbose=verbose and writes, dp=dp)["history"]


def fit(config, dataset, data_loader, val_loader, max_epoch, device, train_writer=None, checkpoint_directory=None,
        verbose=True, model_seed=None, dp=None):
    """The training loop shared by training_script (reference src/train.py:52-252) and the HP-search trainer
    (reference src/parameter_search/hp_model_training.py:13-228): encoder, model, optimiser, loss, regulariser, the
    per-epoch LambdaLR decay, grid-order batches, validation every val_epoch.  Returns {'history': [(epoch, psnr,
    ssim)], 'best_psnr', 'best_psnr_ep', 'best_ssim', 'best_ssim_ep', 'model', 'encoder', 'optim'}."""
    in_image_space = config["transform"]
    encoder = Positional_Encoder(config["encoder"], device=device)
    if model_seed is not None:
        torch.manual_seed(model_seed)               # hp_model_training.py:50 seeds between encoder and model
    model = build_model(config)
    model.to(device=device)
    model.train()
    if config["optimizer"] == "Adam":
        optim = FusedAdam(model, lr=config["lr"], betas=(config["beta1"], config["beta2"]), weight_decay=config["weight_decay"])
    else:
        raise NotImplementedError(config["optimizer"])
    loss_fn = build_loss(config)

    reg_type = config["regularization"]["type"]
    regularization = None
    if reg_type not in ("none", "None", None):
        lam = config["regularization"]["strenght"]
        regularization = {"L1": Regularization_L1(lam), "L2": Regularization_L2(lam)}.get(reg_type)
        if regularization is not None:
            optim.reg_l1 = lam if reg_type == "L1" else 0.0
            optim.reg_l2 = lam if reg_type == "L2" else 0.0

    if "pretrain" in config:                    # reference :117-121
        ckpt = torch.load(config["pretrain"], map_location=device)
        model.load_state_dict(ckpt["net"])
        optim.load_state_dict(ckpt["opt"])
        encoder.B = ckpt["enc"]

    bs = config["batch_size"]
    C, H, W, S = dataset.img_shape
    train_ds = data_loader.ds
    gt_image = M.reconstruct(dataset.image.to(device), (C, H, W), in_image_space)

    wire = config["model"] in ("WIRE", "WIRE2D")           # both take raw coordinates (reference networks.py:234, wire2d.py:92)
    enc_ok = (config["encoder"]["embedding"] == ("none" if wire else "gauss")
              or (config["encoder"]["embedding"] == "LogF" and config["model"] in ("SIREN", "FFN")))
    per_coil = bool(config.get("per_coil", False))
    if per_coil:
        bs = H * W                              # one coil per batch, grid order (reference per-coil loader)
    has_mask = train_ds.coords_mask is not None
    # the reference reaches tv_loss only inside `if len(mask_coords) != 0` (:172-174) and views the batch as (H, W, 2)
    use_tv = bool(config.get("use_tv", False)) and has_mask
    # `LSL` means CenterLoss in this entry point and in the HP-search trainer (reference src/train.py:87-88,
    # hp_model_training.py:81-82) -- randperm-based, not a fused-kernel target; the engine's fused log-space loss is what
    # train_kspace_multiscale.py calls LSL (:113-114)
    # the complex-parameter optimiser kernels of WIRE / WIRE2D carry no regulariser term (the reference's penalty on
    # complex tensors goes through abs / complex pow, regularization.py:27,35): those fits keep the autograd path
    fused = (config["loss"] in FUSABLE_LOSSES and config["loss"] != "LSL" and enc_ok
             and (not use_tv or (per_coil and config["loss"] != "HDR"))
             and (regularization is None or not wire))
    trainer = None
    if fused:
        mask = train_ds.coords_mask[:, 0] if has_mask else None
        trainer = FusedTrainer(model, encoder, optim, config["loss"], bs, train_ds.coords, train_ds.image, mask,
                               config.get("loss_opts"), tv=(H, W) if use_tv else None, dp=dp)
    else:
        if dp is not None and dp.world > 1:
            raise NotImplementedError("data-parallel fits need the fused step (fusable model + loss); run unfused "
                                      "configurations as independent fits, one per GPU")
        if verbose:
            print("unfused path: model(x) -> loss -> backward -> optim.step through the engine's autograd face")

    scheduler = LambdaLR(optim, lambda x: 0.2 ** min(x / max_epoch, 1))
    history = []
    best = {"best_psnr": -999999, "best_psnr_ep": 0, "best_ssim": -1, "best_ssim_ep": 0}
    log_iter = config["log_iter"]
    for epoch in range(max_epoch):
        model.train()
        if trainer is not None:
            for it in range(trainer.steps_per_epoch):
                loss_dev = trainer.step()
                if it % log_iter == log_iter - 1:              # the only host sync of the training loop
                    train_loss = trainer.global_loss(loss_dev)
                    (train_writer.log_train if train_writer else (lambda *a, **k: None))(train_loss, epoch * trainer.steps_per_epoch + it + 1)
                    if verbose:
                        print("[Epoch: {}/{}, Iteration: {}] Train loss: {:.4g}".format(epoch + 1, max_epoch, it, train_loss))
        else:
            for it, (coords, gt, dist_to_center, mask_coords) in enumerate(data_loader):
                kcoords = coords.to(device)
                gt = gt.to(device)
                x = encoder.embedding(kcoords)
                out = model(x)
                optim.zero_grad()
                train_loss = 0
                if len(mask_coords) != 0:
                    if config["use_tv"]:
                        train_loss = train_loss + tv_loss(out.view((H, W, 2)))
                    sel = mask_coords.to(device)[:, 0]
                    out, gt = out[sel], gt[sel]
                if config["loss"] in ["HDR", "LSL", "FFL", "tanh"]:      # reference :178-182 (FFL cannot be unpacked there either)
                    loss, _ = loss_fn(out, gt, kcoords)
                    train_loss = train_loss + loss
                else:
                    train_loss = train_loss + 0.5 * loss_fn(out, gt)
                if regularization is not None:
                    train_loss = train_loss + regularization(model.parameters())
                    l1, l2 = optim.reg_l1, optim.reg_l2        # autograd already carries the penalty here
                    optim.reg_l1 = optim.reg_l2 = 0.0
                train_loss.backward()
                optim.step()
                if regularization is not None:
                    optim.reg_l1, optim.reg_l2 = l1, l2
                if it % log_iter == log_iter - 1:
                    (train_writer.log_train if train_writer else (lambda *a, **k: None))(float(train_loss), epoch * len(data_loader) + it + 1)
        if (epoch + 1) % config["val_epoch"] == 0:
            model.eval()
            with torch.no_grad():
                if trainer is not None:
                    flat = trainer.predict(dataset.coords, chunk=bs)
                else:
                    flat = torch.cat([model(encoder.embedding(c.to(device))) for c, _, _, _ in val_loader])
                recon = M.reconstruct(flat, (C, H, W), in_image_space)
                test_psnr = float(M.psnr(gt_image, recon))
                test_ssim = float(M.ssim(gt_image, recon))
            history.append((epoch + 1, test_psnr, test_ssim))
            if test_psnr > best["best_psnr"]:
                best["best_psnr"], best["best_psnr_ep"] = test_psnr, epoch
            if test_ssim > best["best_ssim"]:
                best["best_ssim"], best["best_ssim_ep"] = test_ssim, epoch
            (train_writer.log_test if train_writer else (lambda *a, **k: None))(0.0, test_psnr, test_ssim, epoch + 1)
            if verbose:
                print("[Validation Epoch: {}/{}] Test psnr: {:.4g} | Test ssim: {:.4g}".format(epoch + 1, max_epoch, test_psnr, test_ssim))
        if checkpoint_directory and (epoch + 1) % config["image_save_epoch"] == 0:
            torch.save({"net": model.state_dict(), "enc": encoder.B, "opt": optim.state_dict()},
                       os.path.join(checkpoint_directory, "model_%06d.pt" % (epoch + 1)))
        scheduler.step()
    return {"history": history, **best, "model": model, "encoder": encoder, "optim": optim}


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="src/config/config_image.yaml", help="Path to the config file.")
    parser.add_argument("--data_samples", type=str, default="", help="Path to the config file.")
    parser.add_argument("--output_path", type=str, default=".", help="outputs path")
    parser.add_argument("--parallel", type=str, default="auto", choices=["auto", "dp", "independent"],
                        help="under torchrun: coordinate data-parallel fits, or one independent fit per GPU")
    opts = parser.parse_args()
    config = set_default_configs(get_config(opts.config))
    data_samples = get_config(opts.data_samples)
    rank, world = init_distributed()

    def loaders(sample, slice_no):
        return get_data_loader(data=config["data"], data_root=config["data_root"], set=config["set"],
                               batch_size=config["batch_size"], transform=config["transform"], num_workers=0,
                               sample=sample, slice=slice_no, shuffle=True, full_norm=config["full_norm"],
                               normalization=config["normalization"], undersampling=config["undersampling"],
                               use_dists="no", per_coil=config["per_coil"])

    # (name, file sample, slice): the reference always loads config["sample"] in the multi-sample loop (:304); kept
    if not data_samples:
        fits = [(config["sample"], config["sample"], config["slice"])]
    else:
        fits = [(sample, config["sample"], _slice) for sample, slices in data_samples["samples"].items() for _slice in slices]
    mode, mine = plan_fits(fits, rank, world, opts.parallel)
    dp = DataParallel() if mode == "dp" else None
    if world > 1 and rank == 0:
        print(f"{world} ranks, {len(fits)} fit(s): {mode}")
    for name, file_sample, _slice in mine:
        dataset, data_loader, val_loader = loaders(file_sample, _slice)
        training_script(config, dataset, data_loader, val_loader, name, _slice, dp=dp)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
