"""Batches of src/data/slices.get_data_loader against the UNMODIFIED reference loader (src/models/utils.py:57-141:
MRIDataset* + torch DataLoader(shuffle=False) + collate_inr, per-coil wrapper included) fed the same synthetic k-space.
Live test: runs where /root/reference exists.  Runs in a subprocess because the reference's top-level package names
(`data`, `models`, `undersampling`) are the same as this repo's drop-in packages."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r'''
import contextlib, importlib, io, json, os, sys, types, warnings
ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
import torch
from oracle import ref_shims
from mri_implicit_neural_representations_b200 import synthetic
C, H, W = 3, 32, 40
k = synthetic.fft2c(synthetic.phantom_slice(1234, C, H, W)).numpy()[None]
CASES = [dict(transform=True, use_dists="no", undersampling=None, per_coil=False, normalization="max"),
         dict(transform=False, use_dists="yes", undersampling="grid-2*1", per_coil=False, normalization="coil"),
         dict(transform=False, use_dists="yes", undersampling="grid-2*1", per_coil=True, normalization="max"),
         dict(transform=False, use_dists="no", undersampling="grid-3*2", per_coil=True, normalization="max")]
BS = 1000

def batches(loader):
    out = []
    for b in loader:
        out.append([x.clone() if isinstance(x, torch.Tensor) else x for x in b])
    return out

def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **kw)

# ---- ours first (its packages are then dropped from sys.modules so the reference's same-named ones can load)
sys.path.insert(0, os.path.join(ROOT, "src"))
from data import slices
ours = []
for c in CASES:
    ds, tl, vl = quiet(slices.get_data_loader, "knee", "data", "train", BS, sample=0, slice=0, shape=(C, H, W), **c)
    ours.append((batches(tl), batches(vl)))
sys.path.remove(os.path.join(ROOT, "src"))
for name in [n for n in sys.modules if n == "data" or n.startswith("data.") or n == "undersampling" or n.startswith("undersampling.")
             or n == "models" or n.startswith("models.")]:
    del sys.modules[name]

# ---- the reference
ref_shims._plant_shims()
plt = sys.modules["matplotlib.pyplot"]
plt.imshow = plt.savefig = plt.show = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = plt
try:
    import torchvision.utils  # noqa: F401
except Exception:
    tv, tvu = types.ModuleType("torchvision"), types.ModuleType("torchvision.utils")
    tv.utils = tvu
    sys.modules.update({"torchvision": tv, "torchvision.utils": tvu})
for alias in ("undersampling", "data", "models"):
    pkg = types.ModuleType(alias)
    pkg.__path__ = [os.path.join(ref_shims.REF_SRC, alias)]
    sys.modules[alias] = pkg
mu = importlib.import_module("models.utils")
nd = sys.modules["data.nerp_datasets"]

def _load(self, root, sample):
    self.file_name = "synthetic"
    return k, (H, W, 1)
nd.MRIDataset._MRIDataset__load_files = _load

report = []
for c, (o_train, o_val) in zip(CASES, ours):
    ds, tl, vl = quiet(mu.get_data_loader, "knee", "data", "train", BS, sample=0, slice=0, **c)
    for which, ref_b, our_b in (("train", batches(tl), o_train), ("val", batches(vl), o_val)):
        rec = {"case": {k2: str(v) for k2, v in c.items()}, "loader": which, "n_ref": len(ref_b), "n_ours": len(our_b), "bad": []}
        for i, (rb, ob) in enumerate(zip(ref_b, our_b)):
            rc, ry, rd, rm = rb
            oc, oy, od, om = ob
            if not torch.equal(rc, oc): rec["bad"].append((i, "coords"))
            if rc.shape != oc.shape or float((ry - oy).abs().max()) > 2e-5 * max(float(ry.abs().max()), 1e-30): rec["bad"].append((i, "image"))
            if isinstance(rd, torch.Tensor) != isinstance(od, torch.Tensor): rec["bad"].append((i, "dist presence"))
            elif isinstance(rd, torch.Tensor) and not torch.allclose(rd.reshape(-1), od.reshape(-1), rtol=0, atol=1e-7): rec["bad"].append((i, "dist"))
            if isinstance(rm, torch.Tensor) != isinstance(om, torch.Tensor): rec["bad"].append((i, "mask presence"))
            elif isinstance(rm, torch.Tensor) and not torch.equal(rm, om): rec["bad"].append((i, "mask"))
        report.append(rec)
print("REPORT " + json.dumps(report))
'''


def test_loader_batches_equal_reference():
    res = subprocess.run([sys.executable, "-c", _CHILD, ROOT], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("REPORT ")][-1]
    report = json.loads(line[len("REPORT "):])
    assert len(report) == 8
    for rec in report:
        assert rec["n_ref"] == rec["n_ours"] and rec["n_ref"] > 0, rec
        assert rec["bad"] == [], rec
