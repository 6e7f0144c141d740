"""Partitioned k-space fit: one independent model per ring, ring -> GPU (drop-in for the reference's
src/train_variations/train_clustering.py on the B200 engine).

Reference behaviour kept: k-means ring partition of the full slice (:118-123), `no_models` SIRENs each with its own Adam
and LambdaLR (:54-59,74-77,164-166), per batch and ring the limits are widened by |N(0,0.05)| (:175-176) and only the
ring's rows enter the loss (:179-188), validation writes every ring's prediction into the batch in ring order with the
un-widened limits (:209-225), checkpoints `submodel_<i>_<epoch>.pt` with {'net','enc','opt'} (:262-268).
What changes: the rings are spread over the ranks of a torchrun launch (one ring model per GPU at 4 GPUs), the batches
are device-resident in grid order, the step is the fused masked step, and the only cross-GPU traffic is one
all-reduce of the reconstructed slice per validation.

    python src/train_variations/train_clustering.py --config <yaml>
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 src/train_variations/train_clustering.py --config <yaml>
"""
import argparse
import os
import sys

import numpy as np
import torch
from torch.optim.lr_scheduler import LambdaLR

_SRC = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_SRC)
for _p in (_SRC, _ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from models.networks import FFN, SIREN, Positional_Encoder                                    # noqa: E402
from clustering import partition_kspace                                                       # noqa: E402
from data.slices import get_data_loader                                                       # noqa: E402
from utils import get_config, set_default_configs                                             # noqa: E402
from mri_implicit_neural_representations_b200 import metrics as M                             # noqa: E402
from mri_implicit_neural_representations_b200 import parallel as P                            # noqa: E402
from mri_implicit_neural_representations_b200.trainer import FusedAdam, RingTrainer           # noqa: E402


@torch.no_grad()
def assemble_rings(trainers, coords, dist_to_center, radii, no_models):
    """Validation assembly (reference :209-232): every ring model predicts its rows (un-widened limits, a row on a shared
    edge takes the outer ring's value); this rank fills in the rings it owns, one all-reduce(sum) completes the slice."""
    some = next(iter(trainers.values()))
    partial = torch.zeros(coords.shape[0], some.eng.plan.out_cols, device=coords.device)
    writers = P.ring_writer_masks(dist_to_center, radii, no_models)
    for i, tr in trainers.items():
        rows = torch.nonzero(writers[i]).squeeze(1)
        if rows.numel():
            partial[rows] = tr.predict_rows(rows, coords)
    return P.combine_ring_outputs(partial)


def training_clustering(config, dataset, data_loader, val_loader, output_path=None, verbose=True, rank=None, world=None,
                        jitter_seed=0, record_losses=False):
    """Returns {'history': [(epoch, psnr, ssim)], 'radii', 'models': {ring: module}, 'losses': {ring: last loss},
    'loss_trace': {ring: [loss of every step]} (only with record_losses=True: one host sync per step)}."""
    if not torch.cuda.is_available():
        raise RuntimeError("this engine has no CPU fallback: a CUDA device (B200) is required")
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
    device = torch.device("cuda", torch.cuda.current_device())
    max_epoch = config["max_epoch"]
    in_image_space = config["transform"]
    no_models, no_steps = config["partition"]["no_models"], config["partition"]["no_steps"]
    if config["model"] not in ("SIREN", "FFN"):
        raise NotImplementedError(config["model"])          # the reference builds the model list for SIREN only (:54-59)

    # one encoder for all rings (:50), then the models in ring order -- every rank builds all of them so that model i
    # gets the same initial weights whatever the number of GPUs; only the owned ones go to the device
    encoder = Positional_Encoder(config["encoder"], device=device)
    mine = P.owned_rings(no_models, rank, world)
    models, optims, scheds, trainers = {}, {}, {}, {}
    _, part_radii = partition_kspace(dataset=dataset, no_steps=no_steps, no_parts=no_models, show=False)
    C, H, W, S = dataset.img_shape
    train_ds = data_loader.ds
    dist_all = torch.sqrt(train_ds.coords[:, 1] ** 2 + train_ds.coords[:, 2] ** 2)             # :169
    bs = config["batch_size"]
    for i in range(no_models):
        m = SIREN(config["net"]) if config["model"] == "SIREN" else FFN(config["net"])
        if i not in mine:
            continue
        m.to(device=device)
        m.train()
        models[i] = m
        optims[i] = FusedAdam(m, lr=config["lr"], betas=(config["beta1"], config["beta2"]), weight_decay=config["weight_decay"])
        scheds[i] = LambdaLR(optims[i], lambda x: 0.2 ** min(x / max_epoch, 1))
        trainers[i] = RingTrainer(m, encoder, optims[i], config["loss"], bs, train_ds.coords, train_ds.image, dist_all,
                                  config.get("loss_opts"))
    if verbose and rank == 0:
        print("Kmeans Radial partitioning:")
        print(part_radii / np.sqrt(2))
        print("rings per rank:", {r: P.owned_rings(no_models, r, world) for r in range(world)})

    gt_image = M.reconstruct(dataset.image.to(device), (C, H, W), in_image_space)
    val_dist = torch.sqrt(dataset.coords[:, 1] ** 2 + dataset.coords[:, 2] ** 2).to(device)
    val_coords = dataset.coords.to(device)
    rng = np.random.RandomState(jitter_seed)
    n_it = (len(train_ds) + bs - 1) // bs
    history, last = [], {i: None for i in mine}
    trace = {i: [] for i in mine}
    ckpt_dir = None
    if output_path:
        ckpt_dir = os.path.join(output_path, "checkpoints")
        os.makedirs(ckpt_dir, exist_ok=True)
    for epoch in range(max_epoch):
        for it in range(n_it):
            limits = P.ring_jitter(rng, part_radii, no_models)           # same stream on every rank
            for i in mine:
                loss_dev = trainers[i].step(it, *limits[i])
                if loss_dev is not None:
                    last[i] = loss_dev
                    if record_losses:
                        trace[i].append(float(loss_dev))
            if verbose and it % config["log_iter"] == config["log_iter"] - 1:
                print("[rank {} Epoch: {}/{}, Iteration: {}] ring losses: {}".format(
                    rank, epoch + 1, max_epoch, it, {i: (None if v is None else float(v)) for i, v in last.items()}))
        if (epoch + 1) % config["val_epoch"] == 0:
            recon_flat = assemble_rings(trainers, val_coords, val_dist, part_radii, no_models)
            recon = M.reconstruct(recon_flat, (C, H, W), in_image_space)
            history.append((epoch + 1, float(M.psnr(gt_image, recon)), float(M.ssim(gt_image, recon))))
            if verbose and rank == 0:
                print("[Validation Epoch: {}/{}] Test psnr: {:.4g} | Test ssim: {:.4g}".format(epoch + 1, max_epoch, *history[-1][1:]))
        if ckpt_dir and (epoch + 1) % config["image_save_epoch"] == 0:
            for i in mine:
                torch.save({"net": models[i].state_dict(), "enc": encoder.B, "opt": optims[i].state_dict()},
                           os.path.join(ckpt_dir, "submodel_%d_%06d.pt" % (i, epoch + 1)))
        for i in mine:
            scheds[i].step()
    return {"history": history, "radii": part_radii, "models": models, "trainers": trainers,
            "losses": {i: (None if v is None else float(v)) for i, v in last.items()}, "loss_trace": trace}


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="src/config/config_image.yaml", help="Path to the config file.")
    parser.add_argument("--output_path", type=str, default=".", help="outputs path")
    opts = parser.parse_args()
    config = set_default_configs(get_config(opts.config))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dataset, data_loader, val_loader = get_data_loader(
        data=config["data"], data_root=config["data_root"], set=config["set"], batch_size=config["batch_size"],
        transform=config["transform"], num_workers=0, sample=config["sample"], slice=config["slice"], shuffle=True,
        full_norm=config["full_norm"], normalization=config["normalization"])
    training_clustering(config, dataset, data_loader, val_loader, os.path.join(opts.output_path, "outputs"))
