"""Drop-in for src/utils.py:8-24 and the config loader of src/models/utils.py:25-32."""
import json

import yaml


def set_default_configs(config):
    config.setdefault("per_coil", False)
    config.setdefault("use_tv", False)
    if "regularization" not in config:
        config["regularization"] = {"type": "none"}
    config.setdefault("undersampling", None)
    return config


def get_config(config):
    if not config:
        return None
    with open(config, "r") as f:
        return json.load(f) if config.endswith(".json") else yaml.load(f, Loader=yaml.Loader)
