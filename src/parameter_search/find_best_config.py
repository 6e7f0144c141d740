"""Grid / random hyper-parameter search -- drop-in for the reference's src/parameter_search/find_best_config.py (same
function names, arguments, search-space format and returned dict).

Every candidate is an independent short fit, so the search shards with no data-path collective: when
torch.distributed is initialised, candidate i runs on rank i % world and the (psnr, ssim) pairs are all-gathered.

Two deliberate differences from the reference, both in the bookkeeping and not in any fit:
  * the reference stores `best_config = model_configs`, the SAME dict it keeps mutating, so it always reports the last
    candidate as best (:77-84); here the winning configuration is deep-copied;
  * the reference never appends to `results` (:37,93), so its 'results' list is empty; here it holds
    (hp_config, {'best_psnr', 'best_ssim', ...}) for every candidate."""
import copy
import os
import random
import sys
from itertools import product
from math import log10

_SRC = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

from parameter_search.hp_model_training import hp_training_function   # noqa: E402

ALLOWED_RANDOM_SEARCH_PARAMS = ["log", "int", "float", "item"]


def update_model_config(model_configs, hp_configs):
    """'net.network_width'-style keys address one level of nesting (reference :15-26)."""
    for k, v in hp_configs.items():
        if "." in k:
            a, b = k.split(".")[:2]
            model_configs[a][b] = v
        else:
            model_configs[k] = v
    return model_configs


def _loaders(cfg):
    from data.slices import get_data_loader
    return get_data_loader(data=cfg["data"], data_root=cfg["data_root"], set=cfg["set"], batch_size=cfg["batch_size"],
                           transform=cfg["transform"], num_workers=0, sample=cfg["sample"], slice=cfg["slice"], shuffle=True,
                           full_norm=cfg["full_norm"], normalization=cfg["normalization"], undersampling=cfg["undersampling"],
                           use_dists="no", per_coil=cfg["per_coil"], **({"shape": tuple(cfg["_shape"])} if "_shape" in cfg else {}))


def findBestConfig(model_configs, hp_configs, epochs, device):
    import torch.distributed as dist
    import yaml
    image_directory = model_configs.pop("image_directory", None)
    output_directory = model_configs.pop("output_directory", None)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    local = {}
    for i, hp in enumerate(hp_configs):
        if i % world != rank:
            continue
        print("\nEvaluating Config #{} [of {}]:\n".format(i + 1, len(hp_configs)), hp)
        cfg = update_model_config(copy.deepcopy(model_configs), hp)
        cfg["config_index"] = i + 1
        if output_directory:
            with open(os.path.join(output_directory, "hp_search_config_{}.yaml".format(i + 1)), "w") as f:
                yaml.dump(hp, f, default_flow_style=False)
        dataset, data_loader, val_loader = _loaders(cfg)
        local[i] = (hp_training_function(config=cfg, max_epoch=epochs, image_directory=image_directory, device=device,
                                         dataset=dataset, data_loader=data_loader, val_loader=val_loader), cfg)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, {i: (st, {k: v for k, v in cfg.items() if not k.startswith("_")}) for i, (st, cfg) in local.items()})
        local = {i: v for part in gathered for i, v in part.items()}
    best_psnr, best_config_psnr, best_ssim, best_config_ssim, results = -999999, None, -1, None, []
    for i in range(len(hp_configs)):
        stats, cfg = local[i]
        results.append(stats)
        if stats["best_psnr"] > best_psnr:
            best_psnr, best_config_psnr = stats["best_psnr"], cfg
        if stats["best_ssim"] > best_ssim:
            best_ssim, best_config_ssim = stats["best_ssim"], cfg
    print("\nSearch done. Best Psnr = {}".format(best_psnr))
    print("Best Config PSNR:", best_config_psnr)
    print("\nSearch done. Best SSIM = {}".format(best_ssim))
    print("Best Config SSIM:", best_config_ssim)
    return {"PSNR": {"config": best_config_psnr}, "SSIM": {"config": best_config_ssim}, "results": list(zip(hp_configs, results))}


def grid_configs(grid_search_spaces):
    """All value combinations in itertools.product order over the keys' insertion order (reference :137-152)."""
    spaces = {k: (v.get("values") if isinstance(v, dict) else v) for k, v in grid_search_spaces.items()}
    return [dict(zip(spaces.keys(), inst)) for inst in product(*spaces.values())]


def grid_search(model_class, model_configs, device, dataset=None, image=None, train_loader=None, val_loader=None, epochs=20,
                grid_search_spaces=None):
    print("Running Grid Search method on model: ", model_class)
    if grid_search_spaces is None:
        grid_search_spaces = {"lr": {"values": [0.0001, 0.001, 0.01, 0.1]}}
    return findBestConfig(model_configs=model_configs, hp_configs=grid_configs(grid_search_spaces), epochs=epochs, device=device)


def random_search_spaces_to_config(random_search_spaces):
    """One sample per key; modes log / int / float / item (reference :186-213, same random-module calls)."""
    config = {}
    for key, (rng, mode) in random_search_spaces.items():
        if mode not in ALLOWED_RANDOM_SEARCH_PARAMS:
            print("'{}' is not a valid random sampling mode. Ignoring hyper-param '{}'".format(mode, key))
        elif mode == "log":
            if rng[0] <= 0 or rng[-1] <= 0:
                print("Invalid value encountered for logarithmic sampling of '{}'. Ignoring this hyper param.".format(key))
                continue
            config[key] = 10 ** random.uniform(log10(rng[0]), log10(rng[-1]))
        elif mode == "int":
            config[key] = random.randint(rng[0], rng[-1])
        elif mode == "float":
            config[key] = random.uniform(rng[0], rng[-1])
        elif mode == "item":
            config[key] = random.choice(rng)
    return config


def random_search(model_class, model_configs, device, num_search=20, epochs=20, random_search_spaces=None):
    print("Running Random Search method on model: ", model_class)
    if random_search_spaces is None:
        random_search_spaces = {"lr": {"values": [0.0001, 0.1], "type": "log"}}
    spaces = {k: ((v.get("values"), v.get("type")) if isinstance(v, dict) else tuple(v)) for k, v in random_search_spaces.items()}
    configs = [random_search_spaces_to_config(spaces) for _ in range(num_search)]
    return findBestConfig(model_configs=model_configs, hp_configs=configs, epochs=epochs, device=device)
