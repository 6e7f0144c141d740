"""The data side of the path: src/data/slices.py (normalisation, flattening, coordinate grid, undersampling mask,
distances) against the UNMODIFIED reference datasets (src/data/nerp_datasets.py: MRIDataset, MRIDatasetUndersampling,
MRIDatasetWithDistances) fed with the same synthetic multi-coil k-space through a monkeypatched file loader.
Live test: runs where /root/reference exists."""
import contextlib
import importlib
import io
import os
import sys
import types
import warnings

import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "src")
C, H, W = 3, 32, 40


@pytest.fixture(scope="module")
def ref_nd():
    ref_shims._plant_shims()
    plt = sys.modules["matplotlib.pyplot"]
    plt.imshow = plt.savefig = plt.show = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = plt
    saved = {k: sys.modules.get(k) for k in ("undersampling", "undersampling.undersampler", "undersampling.utils")}
    for k in saved:
        sys.modules.pop(k, None)
    pkg = types.ModuleType("undersampling")
    pkg.__path__ = [os.path.join(ref_shims.REF_SRC, "undersampling")]
    sys.modules["undersampling"] = pkg
    dpkg = types.ModuleType("inr_ref_data")
    dpkg.__path__ = [os.path.join(ref_shims.REF_SRC, "data")]
    sys.modules["inr_ref_data"] = dpkg
    try:
        nd = importlib.import_module("inr_ref_data.nerp_datasets")
    finally:
        for k in ("undersampling", "undersampling.undersampler", "undersampling.utils"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    from mri_implicit_neural_representations_b200 import synthetic
    k = synthetic.fft2c(synthetic.phantom_slice(1234, C, H, W)).numpy()[None]        # [slices, C, H, W] raw k-space
    nd.MRIDataset._MRIDataset__load_files = lambda self, root, sample: (k, (H, W, 1))
    return nd


@pytest.fixture(scope="module")
def ours():
    sys.path.insert(0, SRC)
    from data import slices
    yield slices
    sys.path.remove(SRC)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


@pytest.mark.parametrize("norm", ["max", "coil", "abs_max", "max_std", "stand", "gaussian_blur", "tonemap"])
def test_kspace_normalisations(ref_nd, ours, norm):
    r = _quiet(ref_nd.MRIDataset, data_class="knee", transform=False, sample=0, slice=0, normalization=norm)
    o = _quiet(ours.SliceDataset, "knee", "data", "train", False, 0, 0, False, norm, None, False, (C, H, W))
    assert tuple(r.shape) == tuple(o.shape) == (C, H, W, 2)
    assert torch.equal(r.coords, o.coords)
    assert float((r.image - o.image).abs().max()) <= 2e-5 * float(r.image.abs().max())


def test_image_space(ref_nd, ours):
    r = _quiet(ref_nd.MRIDataset, data_class="knee", transform=True, sample=0, slice=0, normalization="coil")
    o = _quiet(ours.SliceDataset, "knee", "data", "train", True, 0, 0, False, "coil", None, False, (C, H, W))
    assert torch.equal(r.coords, o.coords)
    assert float((r.image - o.image).abs().max()) <= 2e-5 * float(r.image.abs().max())


@pytest.mark.parametrize("us", ["grid-2*1", "grid-3*2", "random_line-0.4"])
def test_undersampled_dataset(ref_nd, ours, us):
    torch.manual_seed(3)
    r = _quiet(ref_nd.MRIDatasetUndersampling, data_class="knee", transform=False, sample=0, slice=0, normalization="max",
               undersampling=us)
    torch.manual_seed(3)
    o = _quiet(ours.SliceDataset, "knee", "data", "train", False, 0, 0, False, "max", us, False, (C, H, W))
    assert torch.equal(r.coords_mask, o.coords_mask)
    assert torch.equal(r.coords, o.coords)
    assert float((r.image - o.image).abs().max()) <= 2e-5
    c_r, y_r, d_r, m_r = r[17]                      # (coords, image, [], coords mask) -- what GridOrderLoader yields in batches
    assert torch.equal(m_r, o.coords_mask[17]) and torch.equal(c_r, o.coords[17]) and d_r == []
    c_o, y_o, d_o, m_o = next(iter(ours.GridOrderLoader(o, 64)))
    assert torch.equal(c_o[17], c_r) and torch.equal(m_o[17], m_r) and d_o == []


def test_distances(ref_nd, ours):
    r = _quiet(ref_nd.MRIDatasetWithDistances, data_class="knee", transform=False, sample=0, slice=0, normalization="max",
               undersampling=None)
    o = _quiet(ours.SliceDataset, "knee", "data", "train", False, 0, 0, False, "max", None, True, (C, H, W))
    assert torch.allclose(r.dist_to_center.reshape(-1), o.dist_to_center.reshape(-1), rtol=0, atol=1e-7)
