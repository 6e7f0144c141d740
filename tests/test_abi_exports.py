"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/inr_b200.h declares,
and the host-only entry points (plan construction, layout arithmetic, argument validation) behave."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def inr():
    from mri_implicit_neural_representations_b200 import build
    build.build()
    import mri_implicit_neural_representations_b200 as m
    return m


def test_library_exports_every_declared_symbol(inr):
    header = open(os.path.join(ROOT, "include", "inr_b200.h")).read()
    declared = set(re.findall(r"\b(inr_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = C.CDLL(inr.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    from mri_implicit_neural_representations_b200 import _lib
    assert declared == set(_lib.EXPORTS)


NET = {"network_input_size": 512, "network_output_size": 2, "network_depth": 4, "network_width": 256}
ENC = {"embedding": "gauss", "scale": 4, "embedding_size": 256, "coordinates_size": 3}


def test_plan_layout_matches_reference_state_dict_order(inr):
    from oracle import inr_oracle as O
    plan = inr.Plan("SIREN", NET, ENC)
    sd = O.siren_init(dict(NET))
    assert plan.n_params == sum(v.numel() for v in sd.values()) == 263426
    off = 0
    for (o, rows, cols, layer, is_bias), (k, v) in zip(plan.tensors, sd.items()):
        assert o == off and rows * cols == v.numel(), k
        assert is_bias == k.endswith("bias")
        off += v.numel()


@pytest.mark.parametrize("short_tiles", [0, 1])
def test_workspace_layout_is_consistent(inr, monkeypatch, short_tiles):
    monkeypatch.setenv("INR_CHAIN_T", str(short_tiles))
    plan = inr.Plan("SIREN", NET, ENC)
    prev = 0
    for bs in (1, 128, 129, 2368, 2369, 10000, 11840, 11841, 102400):
        lay = plan.workspace_layout(bs)
        # below one wave of 80-row tiles (148 SMs without a GPU) the chain is cut into one short tile per SM (chain_t.cu)
        rows = 128 if (bs > 148 * 80 or not short_tiles) else max(16, 16 * -(-bs // (148 * 16)))
        assert lay["tile_rows"] == rows and lay["lb"] == (2048 if rows == 128 else rows * 16 + 16)
        assert lay["n_tiles"] == -(-bs // rows)
        # a workspace sized for bs rows serves every smaller batch: never smaller than the layout, and monotone
        assert plan.workspace_bytes(bs) >= lay["total"] and plan.workspace_bytes(bs) >= prev
        prev = plan.workspace_bytes(bs)
        regions = sorted([lay["scal"], lay["part"], lay["g"], *lay["h"][:4], *lay["d"][:3], *lay["dz"][:3],
                          lay["dzlast"], lay["gpart"]])
        assert len(set(regions)) == len(regions) and regions[-1] < lay["total"]
        assert 1 <= lay["n_split"] <= lay["n_tiles"]


def test_wgrad_split_schedule(inr):
    """The split-K factor comes from the two-class wgrad schedule (abi.cu: build_wgrad_sched / wgrad_splits): heavy
    (unit, split) items one per CTA, light items two per CTA, everything inside one wave of the device's SMs; models with
    more units than SMs keep split factor 1.  Pinned for the BASELINE shapes on a 148-SM device (the default without a GPU)."""
    import torch
    if torch.cuda.is_available() and torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("schedule figures are for 148 SMs")
    wire = inr.Plan("WIRE", {"network_input_size": 3, "network_output_size": 2, "network_depth": 4, "network_width": 256,
                             "first_omega_0": 30, "hidden_omega_0": 30, "scale": 15}, {"embedding": "none"})
    assert wire.workspace_layout(25000)["n_split"] == 10        # 12 heavy x 10 + ceil(5 light x 10 / 2) = 145 CTAs
    assert wire.workspace_layout(300)["n_split"] == 3           # never more splits than row tiles
    siren = inr.Plan("SIREN", NET, ENC)
    assert siren.workspace_layout(10000)["n_split"] == 16       # 8 heavy x 16 + ceil(2 light x 16 / 2) = 144 CTAs
    gabor = inr.Plan("Gabor", {"network_input_size": 512, "network_output_size": 2, "network_depth": 8, "network_width": 512}, ENC)
    assert gabor.workspace_layout(102400)["n_split"] == 1


def test_unsupported_shapes_fail_loudly(inr):
    with pytest.raises(inr.InrError):
        inr.Plan("SIREN", dict(NET, network_width=192), ENC)
    with pytest.raises(inr.InrError):
        inr.Plan("SIREN", dict(NET, network_input_size=500), {"embedding": "none"})
    with pytest.raises(NotImplementedError):
        inr.Plan("NoSuchModel", NET, ENC)


def test_no_cpu_fallback(inr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    plan = inr.Plan("SIREN", NET, ENC)
    with pytest.raises(inr.InrError):
        inr.ChainEngine(plan, max_batch=256)
