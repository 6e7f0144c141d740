"""Ring partition of k-space (src/clustering.py = restatement of reference src/clustering.py:19-134).

Pinned (a) live against the unmodified reference functions where /root/reference exists, (b) by the committed
fixture tests/golden/clustering.json (generated from the reference by oracle/make_golden_clustering.py)."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden_clustering import synthetic_kspace_dataset, CASES      # noqa: E402


def _ours():
    import importlib.util
    spec = importlib.util.spec_from_file_location("inr_src_clustering", os.path.join(ROOT, "src", "clustering.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", sorted(CASES))
def test_partition_matches_golden(case):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "clustering.json")))[case]
    ds = synthetic_kspace_dataset(**CASES[case]["data"])
    Cl = _ours()
    labels, radii = Cl.partition_kspace(dataset=ds, show=False, **CASES[case]["part"])
    assert list(map(int, labels)) == gold["labels"]
    assert np.array_equal(np.asarray(radii, dtype=np.float64), np.asarray(gold["radii"]))
    stats, radii2 = Cl.partition_and_stats(dataset=ds, show=False, stat="max", **CASES[case]["part"])
    assert np.array_equal(np.asarray(radii2), np.asarray(gold["radii"]))
    assert [float(s) for s in stats] == gold["stats_max"]
    ring = Cl.ring_log_max(ds.image, torch.sqrt(ds.coords[:, 1] ** 2 + ds.coords[:, 2] ** 2), CASES[case]["part"]["no_steps"])
    assert [float(v) for v in ring] == gold["ring_log_max"]


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present")
@pytest.mark.parametrize("case", sorted(CASES))
def test_partition_matches_reference_live(case):
    from make_golden_clustering import load_reference_clustering
    R = load_reference_clustering()
    ds = synthetic_kspace_dataset(**CASES[case]["data"])
    Cl = _ours()
    lab_r, rad_r = R.partition_kspace(dataset=ds, show=False, **CASES[case]["part"])
    lab_o, rad_o = Cl.partition_kspace(dataset=ds, show=False, **CASES[case]["part"])
    assert np.array_equal(lab_r, lab_o) and np.array_equal(rad_r, rad_o)
    for stat in ("max", "min"):
        st_r, _ = R.partition_and_stats(dataset=ds, show=False, stat=stat, **CASES[case]["part"])
        st_o, _ = Cl.partition_and_stats(dataset=ds, show=False, stat=stat, **CASES[case]["part"])
        assert torch.equal(st_r, st_o)


def test_edge_points_count_for_both_rings():
    Cl = _ours()
    # distances exactly on ring edges (float32-rounded like torch's comparison) and strictly inside
    edges = Cl.ring_edges(4)
    d = torch.tensor([0.0, edges[0][1], edges[1][1], 0.9, 1.4], dtype=torch.float32)
    img = torch.tensor([[1.0, 0], [5.0, 0], [3.0, 0], [2.0, 0], [7.0, 0]])
    got = Cl.ring_log_max(img, d, 4)
    want = []
    for r0, r1 in edges:
        sel = (d >= r0) & (d <= r1)
        want.append(float(torch.log(img[sel].pow(2).sum(-1).sqrt()).max()))
    assert [float(v) for v in got] == want
