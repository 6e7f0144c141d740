"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/undersampler.json from the UNMODIFIED reference Undersampler
(src/undersampling/undersampler.py) with fixed seeds: per case the mask's SHA-256, its number of kept samples and the
first kept positions.

    python oracle/make_golden_undersampler.py        (needs /root/reference; run in the build container)
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "grid_2x1": ("grid", [2, 1], (3, 12, 10, 2)),
    "grid_3x3": ("grid", [3, 3], (2, 12, 12, 2)),
    "random_line_0.3": ("random_line", [0.3], (2, 16, 20, 2)),
    "radial_2": ("radial", [2], (2, 32, 32, 2)),
    "radial_4_odd": ("radial", [4], (3, 33, 40, 2)),
}
TORCH_SEED, NP_SEED = 5, 7


def digest(mask: torch.Tensor):
    m = mask[:, 0].to(torch.uint8).numpy()
    return {"sha256": hashlib.sha256(m.tobytes()).hexdigest(), "kept": int(m.sum()), "n": int(m.size),
            "first_kept": [int(i) for i in np.flatnonzero(m)[:8]]}


def run(undersampler_cls, set_rng):
    out = {}
    for name, (method, params, shape) in CASES.items():
        img = torch.rand(shape, generator=torch.Generator().manual_seed(1))
        torch.manual_seed(TORCH_SEED)
        u = undersampler_cls(method)
        set_rng(u)
        _, _, grid_mask = u.apply(img, list(params))
        out[name] = digest(grid_mask)
    return out


def main():
    from oracle import ref_shims
    ref_shims._plant_shims()
    plt = sys.modules["matplotlib.pyplot"]
    plt.imshow = plt.savefig = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = plt
    ref_src = ref_shims.REF_SRC

    def load(path, name):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    utils = load(os.path.join(ref_src, "undersampling", "utils.py"), "undersampling.utils")
    pkg = types.ModuleType("undersampling")
    pkg.utils = utils
    sys.modules.update({"undersampling": pkg, "undersampling.utils": utils})
    R = load(os.path.join(ref_src, "undersampling", "undersampler.py"), "inr_reference.undersampler")
    real_rs = np.random.RandomState
    np.random.RandomState = lambda *a, **k: real_rs(NP_SEED)       # the reference's radial mask draws from an unseeded RandomState
    try:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            out = run(R.Undersampler, lambda u: None)
    finally:
        np.random.RandomState = real_rs
    path = os.path.join(ROOT, "tests", "golden", "undersampler.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
