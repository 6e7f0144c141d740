"""HP-search trainer -- drop-in for the reference's src/parameter_search/hp_model_training.py:13-228 (same signature and
returned dict).  The fit itself is the engine's training loop (train.fit): fused CUDA-graph steps where the config is
fusable, the autograd face otherwise.  Like the reference it seeds the model initialisation with 42 AFTER the encoder
has drawn its matrix (:47-50), constructs WIRE2D on request (:55-56), does not write checkpoints or TensorBoard logs.
The PNG dumps of the reference (:33-40,188-214) are not reproduced."""
import os
import sys

_SRC = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _SRC not in sys.path:
    sys.path.insert(0, _SRC)

import train as _train      # noqa: E402


def hp_training_function(config, max_epoch, image_directory, device, dataset=None, data_loader=None, val_loader=None,
                         verbose=False):
    cfg = dict(config)
    cfg["_allow_wire2d"] = True
    res = _train.fit(cfg, dataset, data_loader, val_loader, max_epoch, device, verbose=verbose, model_seed=42)
    return {k: res[k] for k in ("best_psnr", "best_psnr_ep", "best_ssim", "best_ssim_ep")}
