"""Hyper-parameter search entry point -- drop-in for the reference's src/hp_search_script.py (same CLI: --config,
--hp_config, --output_path; same output files best_psnr_config.yaml, best_ssim_config.yaml, configs_and_results.txt).
Launch under torchrun to spread the candidates over the GPUs of a node (independent fits, no data-path collective)."""
import argparse
import os
import shutil
import sys
from datetime import datetime

import yaml

_SRC = os.path.dirname(os.path.abspath(__file__))
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

from parameter_search.find_best_config import grid_search, random_search   # noqa: E402
from train import get_device                                               # noqa: E402
from utils import get_config, set_default_configs                          # noqa: E402


def run(config, hp_config, opts):
    import torch
    import torch.distributed as dist
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    device = get_device(config["model"])
    rank = dist.get_rank() if dist.is_initialized() else 0
    output_folder = os.path.splitext(os.path.basename(opts.config))[0]
    model_name = os.path.join(output_folder, config["data"] + "/img_{}_{}_{}_{}_{}_lr{:.2g}_encoder_{}_hp_{}_search_".format(
        config["model"], config["net"]["network_input_size"], config["net"]["network_width"], config["net"]["network_depth"],
        config["loss"], config["lr"], config["encoder"]["embedding"], hp_config["method"]))
    if config["encoder"]["embedding"] != "none":
        model_name += "_scale{}_size{}".format(config["encoder"]["scale"], config["encoder"]["embedding_size"])
    output_directory = os.path.join(opts.output_path + "/outputs", model_name + datetime.now().strftime("%Y-%m-%d_%H-%M"))
    image_directory = os.path.join(output_directory, "images")
    os.makedirs(image_directory, exist_ok=True)
    if rank == 0:
        shutil.copy(opts.config, os.path.join(output_directory, "config.yaml"))
    config["image_directory"], config["output_directory"] = image_directory, output_directory
    method = hp_config.pop("method")
    if method == "grid":
        best = grid_search(model_configs=config, model_class=config["model"], epochs=hp_config.pop("max_epoch"),
                           grid_search_spaces=hp_config.pop("search_space"), device=device)
    else:
        best = random_search(model_configs=config, model_class=config["model"], num_search=hp_config.pop("num_search"),
                             epochs=hp_config.pop("max_epoch"), random_search_spaces=hp_config.pop("search_space"), device=device)
    if rank == 0:
        for key, name in (("PSNR", "best_psnr_config.yaml"), ("SSIM", "best_ssim_config.yaml")):
            with open(os.path.join(output_directory, name), "w") as f:
                yaml.dump({k: v for k, v in best[key]["config"].items() if not str(k).startswith("_")}, f, default_flow_style=False)
        with open(os.path.join(output_directory, "configs_and_results.txt"), "w") as tf:
            for item in best["results"]:
                tf.write("{} -> {}\n".format(item[0], item[1]))
    return best


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="src/config/config_image.yaml", help="Path to the config file.")
    parser.add_argument("--hp_config", type=str, default="src/hp_config/config_image.json", help="Path to the HP config file.")
    parser.add_argument("--output_path", type=str, default=".", help="outputs path")
    opts = parser.parse_args()
    run(set_default_configs(get_config(opts.config)), get_config(opts.hp_config), opts)
