"""Undersampling masks (src/undersampling/undersampler.py) against the unmodified reference (live, where /root/reference
exists: same seeds -> identical masks, grids and masked images) and the reference's own unit-test properties
(src/tests/undersampler_test.py:45-140), which also run on the GPU box."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ours():
    return _load(os.path.join(ROOT, "src", "undersampling", "undersampler.py"), "inr_src_undersampler")


def _reference():
    """reference undersampler with matplotlib stubbed (it saves a PNG of every mask) and its utils under a private name."""
    ref_shims._plant_shims()
    plt = sys.modules["matplotlib.pyplot"]
    plt.imshow = lambda *a, **k: None
    plt.savefig = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = plt
    saved = {k: sys.modules.get(k) for k in ("undersampling", "undersampling.utils")}
    utils = _load(os.path.join(ref_shims.REF_SRC, "undersampling", "utils.py"), "undersampling.utils")
    pkg = types.ModuleType("undersampling")
    pkg.utils = utils
    sys.modules.update({"undersampling": pkg, "undersampling.utils": utils})
    try:
        mod = _load(os.path.join(ref_shims.REF_SRC, "undersampling", "undersampler.py"), "inr_reference.undersampler")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, utils


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
@pytest.mark.parametrize("method,params,shape", [("grid", [2, 1], (3, 12, 10, 2)), ("grid", [3, 3], (2, 12, 12, 2)),
                                                 ("random_line", [0.3], (2, 16, 20, 2)), ("radial", [2], (2, 32, 32, 2)),
                                                 ("radial", [4], (3, 33, 40, 2)), ("radial", [3], (2, 40, 31, 2))])
def test_masks_equal_reference(ours, method, params, shape, capsys):
    R, _ = _reference()
    img = torch.rand(shape, generator=torch.Generator().manual_seed(1))
    real_rs = np.random.RandomState
    try:
        np.random.RandomState = lambda *a, **k: real_rs(7)       # the reference's radial mask draws from an unseeded RandomState
        torch.manual_seed(5)
        ru = R.Undersampler(method)
        r_img, r_grid, r_mask = ru.apply(img, list(params))
    finally:
        np.random.RandomState = real_rs
    torch.manual_seed(5)
    ou = ours.Undersampler(method)
    ou.rng = real_rs(7)
    o_img, o_grid, o_mask = ou.apply(img, list(params))
    assert torch.equal(r_mask, o_mask)
    assert torch.equal(r_grid, o_grid)
    assert torch.equal(r_img, o_img)
    assert 0 < int(o_mask.sum()) <= o_mask.numel()


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
def test_square_perimeter_closed_form_equals_reference_list(ours):
    _, U = _reference()
    for side in (2, 4, 6, 10):
        for sid in range(side // 2):
            want = U.get_square_ordered_idxs(side, sid)
            r, c = ours.square_perimeter_points(side, sid, np.arange(len(want)))
            assert [(int(a), int(b)) for a, b in zip(r, c)] == [tuple(w) for w in want]


def test_reference_unit_test_properties(ours):
    """the assertions of reference src/tests/undersampler_test.py:45-140"""
    img = torch.rand((3, 12, 12, 2))
    u_img, grid, gmask = ours.Undersampler("grid").apply(img, [3, 3])
    assert u_img.shape[0] * u_img.shape[1] * u_img.shape[2] == grid.shape[0]
    assert int(gmask.sum()) // 3 == grid.shape[0] // 9
    img = torch.rand((2, 64, 64, 3))
    u_img, grid, gmask = ours.Undersampler("random_line").apply(img, [1.0])
    assert u_img.shape[0] * u_img.shape[1] * u_img.shape[2] == grid.shape[0]
    assert int(gmask.sum()) // 3 == grid.shape[0]
    u_img, grid, gmask = ours.Undersampler("radial").apply(img, [2])
    assert u_img.shape[0] * u_img.shape[1] * u_img.shape[2] == grid.shape[0]
    acc = gmask.numel() / int(gmask.sum())
    assert 1.2 < acc < 4.0                         # roughly the requested acceleration of 2
    with pytest.raises(AssertionError):
        ours.Undersampler("spiral")
    assert ours.parse_undersampling_argument("grid-2*1") == ("grid", [2, 1])
    assert ours.parse_undersampling_argument("radial-4") == ("radial", [4.0])
    assert ours.parse_undersampling_argument("random_line-0.5") == ("random_line", [0.5])
    assert ours.parse_undersampling_argument("none") == ("none", [])
    with pytest.raises(ValueError):
        ours.parse_undersampling_argument("spiral-3")


def test_masks_match_golden_from_reference(ours):
    """tests/golden/undersampler.json was generated from the unmodified reference (oracle/make_golden_undersampler.py);
    this pin also holds where the reference tree is absent."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_undersampler as MG
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "undersampler.json")))

    def set_rng(u):
        u.rng = np.random.RandomState(MG.NP_SEED)
    got = MG.run(ours.Undersampler, set_rng)
    assert got == gold
