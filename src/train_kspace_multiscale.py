"""Multi-scale k-space training -- drop-in for the reference's src/train_kspace_multiscale.py (same CLI and config keys).

The whole loop body of the reference (:164-192) -- model(coords, dist_to_center), per-head loss on the FULL target
(`limit_kspace` is a no-op there, :34-39), 0.1 * ConsistencyLoss on all rows, backward, Adam -- runs as ONE fused CUDA-graph
step (FusedTrainer -> inr_train_step_dist) for L2 / L1 / MSLE / LSL without TV.  HDR / tanh per-head losses and the TV term
keep the reference's loop shape on the module's autograd face (composite loss in PyTorch on the [bs, 2] head outputs)."""
import argparse
import os
import sys
from datetime import datetime

import torch
from torch.optim.lr_scheduler import LambdaLR

_SRC = os.path.dirname(os.path.abspath(__file__))
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

from models.networks import Positional_Encoder                               # noqa: E402
from models.mfn import MultiscaleKFourier, MultiscaleBoundedFourier           # noqa: E402
from metrics.losses import ConsistencyLoss, HDRLoss_FF, LogSpaceLoss, MSLELoss, TanhL2Loss, tv_loss  # noqa: E402
from data.slices import get_data_loader                                      # noqa: E402
from clustering import partition_and_stats                                   # noqa: E402
from utils import get_config, set_default_configs                            # noqa: E402
from mri_implicit_neural_representations_b200 import metrics as M            # noqa: E402
from mri_implicit_neural_representations_b200.trainer import FusedAdam, FusedTrainer    # noqa: E402
from log_handler.logger import INRLogger                                      # noqa: E402


def create_pairs(values, multiplication_factor=1):
    """Discs (values[0], values[i+1]) of growing radius, each repeated (reference src/train_kspace_multiscale.py:42-47)."""
    pairs = [(values[0], values[i + 1]) for i in range(len(values) - 1)]
    return [(p[0], p[1]) for p in pairs for _ in range(multiplication_factor)]


def training_multiscale(config, dataset, data_loader, val_loader, output_path=".", verbose=True):
    if not torch.cuda.is_available():
        raise RuntimeError("this engine has no CPU fallback: a CUDA device (B200) is required")
    device = torch.device("cuda")
    max_epoch = config["max_epoch"]
    C, H, W, S = dataset.img_shape
    # ring partition of the FULL slice (reference :72-86): k-means over per-ring max log|k| -> partition radii
    part_config = config["partition"]
    part_mx, part_radii = partition_and_stats(dataset=dataset, no_steps=part_config["no_steps"], no_parts=part_config["no_models"],
                                              stat="max", show=False)
    part_mx = torch.cat((torch.as_tensor(part_mx, dtype=torch.float32).reshape(-1), torch.ones(1)))      # reference :81
    pairs = [(float(a), float(b)) for a, b in create_pairs(part_radii, 1)]
    encoder = Positional_Encoder(config["encoder"], device=device)
    if config["model"] == "Fourier":
        model = MultiscaleKFourier(config["net"])
    elif config["model"] == "BoundedFourier":
        pairs_model = [(float(a), float(b)) for a, b in create_pairs(part_radii, 2)]     # reference :85,:96 -> 8 BoundedLinears
        model = MultiscaleBoundedFourier(config["net"], boundaries=pairs_model)
    else:
        raise NotImplementedError(config["model"])
    model.to(device)
    model.train()
    optim = FusedAdam(model, lr=config["lr"], betas=(config["beta1"], config["beta2"]), weight_decay=config["weight_decay"])
    loss_name = config["loss"]
    if loss_name == "L2":
        loss_fn = torch.nn.MSELoss()
    elif loss_name == "LSL":
        loss_fn = LogSpaceLoss(config["loss_opts"])
    elif loss_name == "MSLE":
        loss_fn = MSLELoss()
    elif loss_name == "HDR":
        loss_fn = HDRLoss_FF(config["loss_opts"])
    elif loss_name == "tanh":
        loss_fn = TanhL2Loss()
    else:
        raise NotImplementedError(loss_name)
    consistency = ConsistencyLoss(pairs)
    scheduler = LambdaLR(optim, lambda x: 0.2 ** min(x / max_epoch, 1))
    gt_image = M.reconstruct(dataset.image.to(device), (C, H, W), config["transform"])
    train_ds = data_loader.ds
    has_mask = train_ds.coords_mask is not None
    use_tv = bool(config.get("use_tv", False))
    # the per-partition maxima weight the HDR / tanh branch (reference :74-81, :185-187)
    mx = [float(v) for v in part_mx]
    run_name = "{}_{}_{}".format(config["model"], loss_name, datetime.now().strftime("%Y-%m-%d_%H-%M-%S"))
    train_writer = INRLogger(os.path.join(output_path, "logs", run_name))
    checkpoint_directory = os.path.join(output_path, "outputs", run_name, "checkpoints")
    os.makedirs(checkpoint_directory, exist_ok=True)
    fused = loss_name in ("L2", "L1", "MSLE", "LSL") and not use_tv and config["encoder"]["embedding"] == "gauss"
    trainer = None
    if fused:
        bs = H * W if config.get("per_coil", False) else config["batch_size"]
        trainer = FusedTrainer(model, encoder, optim, loss_name, bs, train_ds.coords, train_ds.image,
                               train_ds.coords_mask[:, 0] if has_mask else None, config.get("loss_opts"),
                               dist=train_ds.dist_to_center, consistency=(pairs, 0.1))
    elif verbose:
        print("unfused path: model(x, dist) -> composite loss -> backward -> optim.step through the engine's autograd face")
    history = []
    log_iter = config["log_iter"]
    for epoch in range(max_epoch):
        if trainer is not None:
            for it in range(trainer.steps_per_epoch):
                loss = trainer.step()
                if it % log_iter == log_iter - 1:          # the only host sync of the training loop
                    train_writer.log_train(float(loss), epoch * trainer.steps_per_epoch + it + 1)
        else:
            for it, (coords, gt, dist, mask_coords) in enumerate(data_loader):
                kcoords = coords.to(device)
                gt = gt.to(device)
                dist = dist.to(device)
                outs = model(coords=encoder.embedding(kcoords), dist_to_center=dist)
                optim.zero_grad()
                loss = 0
                if use_tv:                                  # reference :174-175: gated on use_tv only
                    loss = loss + tv_loss(outs[-1].view((H, W, 2)))
                sel = None
                if len(mask_coords) != 0:
                    sel = mask_coords.to(device)[:, 0]
                    gt = gt[sel]
                loss = loss + 0.1 * consistency(outs, dist)     # all rows, before the mask (reference :179)
                for idx, o in enumerate(outs):              # every head is supervised on the FULL target (reference :34-39 no-op)
                    if sel is not None:
                        o = o[sel]
                    if loss_name in ("HDR", "tanh"):
                        l, _ = loss_fn(o, gt, gt)           # reference :184 hands the targets in as `kcoords`
                        loss = loss + l / mx[idx]
                    else:
                        loss = loss + 0.5 * loss_fn(o, gt)
                loss.backward()
                optim.step()
                if it % log_iter == log_iter - 1:
                    train_writer.log_train(float(loss), epoch * len(data_loader) + it + 1)
        if (epoch + 1) % config["val_epoch"] == 0:
            model.eval()
            with torch.no_grad():
                if trainer is not None:
                    flat = trainer.predict(dataset.coords)[:, -2:].contiguous()                   # last head (reference :207-212)
                else:
                    flat = torch.cat([model(coords=encoder.embedding(c.to(device)),
                                            dist_to_center=(d.to(device) if len(d) != 0 else None))[-1]
                                      for c, _, d, _ in val_loader])
                recon = M.reconstruct(flat, (C, H, W), config["transform"])
                history.append((epoch + 1, float(loss), float(M.psnr(gt_image, recon)), float(M.ssim(gt_image, recon))))
            if verbose:
                print("[Validation Epoch: {}/{}] loss {:.4g} psnr {:.4g} ssim {:.4g}".format(epoch + 1, max_epoch, *history[-1][1:]))
            train_writer.log_test(0.0, history[-1][2], history[-1][3], epoch + 1)
            model.train()
        if (epoch + 1) % config["image_save_epoch"] == 0:          # reference :239-243
            torch.save({"net": model.state_dict(), "enc": encoder.B, "opt": optim.state_dict()},
                       os.path.join(checkpoint_directory, "model_%06d.pt" % (epoch + 1)))
        scheduler.step()
    return history


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="src/config/config_image.yaml", help="Path to the config file.")
    parser.add_argument("--output_path", type=str, default=".", help="outputs path")
    opts = parser.parse_args()
    config = set_default_configs(get_config(opts.config))
    dataset, data_loader, val_loader = get_data_loader(
        data=config["data"], data_root=config["data_root"], set=config["set"], batch_size=config["batch_size"],
        transform=config["transform"], num_workers=0, sample=config["sample"], slice=config["slice"], shuffle=True,
        full_norm=config["full_norm"], normalization=config["normalization"], undersampling=config["undersampling"],
        use_dists="yes", per_coil=config["per_coil"])
    training_multiscale(config, dataset, data_loader, val_loader, opts.output_path)
