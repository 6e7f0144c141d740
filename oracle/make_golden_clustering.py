"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/clustering.json from the UNMODIFIED reference
(src/clustering.py: partition_kspace / partition_and_stats) on seeded synthetic multi-coil k-space slices.

    python oracle/make_golden_clustering.py        (needs /root/reference; run in the build container)
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "knee_6x96x96_40_4": {"data": dict(seed=1234, C=6, H=96, W=96), "part": dict(no_steps=40, no_parts=4)},
    "knee_3x64x80_20_3": {"data": dict(seed=77, C=3, H=64, W=80), "part": dict(no_steps=20, no_parts=3)},
}


class _DS:
    pass


def synthetic_kspace_dataset(seed, C, H, W):
    """Dataset-like object with the attributes the reference reads: shape (C,H,W,2), image [N,2], coords [N,3]."""
    from mri_implicit_neural_representations_b200 import synthetic
    img = synthetic.phantom_slice(seed, C, H, W)
    k = torch.view_as_real(synthetic.fft2c(img))
    k = k / k.abs().max()                                   # normalization "max"
    ds = _DS()
    ds.shape = (C, H, W, 2)
    ds.image = k.reshape(-1, 2).float().contiguous()
    ds.coords = synthetic.coords_grid(C, H, W).float().contiguous()
    return ds


def load_reference_clustering():
    """Import reference src/clustering.py; its `from models.utils import ...` line (unused by the partition
    functions) is satisfied by a throw-away stub that is removed again so the repo's own src/models is unaffected."""
    from oracle import ref_shims
    ref_shims._plant_shims()
    saved = {k: sys.modules.get(k) for k in ("models", "models.utils")}
    pkg, ut = types.ModuleType("models"), types.ModuleType("models.utils")
    ut.get_config = ut.get_data_loader = None
    pkg.utils = ut
    sys.modules.update({"models": pkg, "models.utils": ut})
    try:
        mod = ref_shims.load("clustering")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def main():
    R = load_reference_clustering()
    out = {}
    for name, c in CASES.items():
        ds = synthetic_kspace_dataset(**c["data"])
        labels, radii = R.partition_kspace(dataset=ds, show=False, **c["part"])
        stats, _ = R.partition_and_stats(dataset=ds, show=False, stat="max", **c["part"])
        dist = torch.sqrt(ds.coords[:, 1] ** 2 + ds.coords[:, 2] ** 2)
        import math
        ring = []
        for i in range(c["part"]["no_steps"]):                # the reference's own ring loop (:48-61)
            r0 = 0 if i == 0 else math.sqrt(2) * i / c["part"]["no_steps"]
            r1 = math.sqrt(2) if i == c["part"]["no_steps"] - 1 else math.sqrt(2) * (i + 1) / c["part"]["no_steps"]
            sel = (dist >= r0) & (dist <= r1)
            ring.append(float(torch.log(ds.image[sel].pow(2).sum(-1).sqrt()).max()))
        out[name] = {"labels": [int(v) for v in labels], "radii": [float(v) for v in radii],
                     "stats_max": [float(v) for v in stats], "ring_log_max": ring}
    path = os.path.join(ROOT, "tests", "golden", "clustering.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
