"""The oracle restatement reproduces the committed reference digests (tests/golden/*.json).

Runs anywhere (no GPU, no /root/reference).  Tolerances: the oracle uses the same fp32 torch ops
in a different association in places (closed-form loss gradients, re/im-split Gabor), so exact
equality is not expected; 2e-4 relative on raw values after 3 Adam steps is what fp32 allows
for the chaotic WIRE stack, 2e-5 for everything else."""
import pytest

from oracle import golden_util as G
from oracle.cases import run_oracle_case

TOL = {"wire_hdr": 5e-3, "wire_l2": 2e-3}
# WIRE is chaotic in fp32 (omega 30, sigma 15, six Gabor layers; SURVEY.md section 7 hard part 1):
# re-associating the same fp32 maths moves first-layer gradients by ~1e-2, more through HDR's
# 1/|e| gradient.  The HDR formula itself is pinned tightly in test_loss_closed_forms below.
GRAD_TOL = {"wire_hdr": 8e-2, "wire_l2": 2e-2}


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_matches_golden(name):
    gold = G.load_golden(name)
    mine = run_oracle_case(name)
    tol = TOL.get(name, 5e-5)
    # initial parameters and encoder matrix: bit-identical RNG restatement
    for k, dg in gold["init"].items():
        assert mine["init"][k]["sum"] == dg["sum"], k
        assert mine["init"][k]["head"] == dg["head"], k
    if gold["encB"] is not None:
        assert mine["encB"]["head"] == gold["encB"]["head"]
    assert abs(mine["loss"] - gold["loss"]) <= tol * abs(gold["loss"])
    assert not G.digest_close(gold["out"], mine["out"], tol)
    assert set(mine["grads"]) == set(gold["grads"])
    for k in gold["grads"]:
        fails = G.digest_close(gold["grads"][k], mine["grads"][k], GRAD_TOL.get(name, 2e-4), atol_scale=1.0)
        assert not fails, (k, fails)
    if name.startswith("wire"):
        return      # fp32 WIRE trajectories decorrelate within 2-3 Adam steps (sign-like first
                    # updates on noisy gradients); the fp64 test below pins the multi-step maths
    for a, b in zip(gold["losses"], mine["losses"]):
        assert abs(a - b) <= GRAD_TOL.get(name, 1e-4) * abs(a)
    for k in gold["final"]:
        fails = G.digest_close(gold["final"][k], mine["final"][k], GRAD_TOL.get(name, 1e-4))
        assert not fails, (k, fails)


import torch
from oracle import inr_oracle as O


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_fp64_matches_reference_fp64(name):
    """Well-conditioned pin: in float64 the oracle's formulas and the reference's modules agree to
    1e-7 on everything, including three Adam steps of the chaotic WIRE+HDR case."""
    gold = G.load_golden(name)["fp64"]
    mine = run_oracle_case(name, torch.float64)
    assert abs(mine["loss"] - gold["loss"]) <= 1e-9 * abs(gold["loss"])
    assert not G.digest_close(gold["out"], mine["out"], 1e-8)
    for k in gold["grads"]:
        fails = G.digest_close(gold["grads"][k], mine["grads"][k], 1e-7)
        assert not fails, (k, fails)
    for a, b in zip(gold["losses"], mine["losses"]):
        assert abs(a - b) <= 1e-6 * abs(a)
    for k in gold["final"]:
        fails = G.digest_close(gold["final"][k], mine["final"][k], 1e-6)
        assert not fails, (k, fails)


@pytest.mark.parametrize("kind", list(G.LOSS_CASES))
def test_loss_closed_forms(kind):
    """Closed-form loss value and d(loss)/d(out) (what the CUDA loss kernel implements) against the
    reference loss classes + autograd, digests in tests/golden/losses.json."""
    gold = G.load_golden("losses")[kind]
    opts = G.LOSS_CASES[kind]
    out, gt, kc, extra, dist = G.loss_case_inputs(kind)
    if kind == "Consistency":
        val, grads = O.loss_consistency([out] + extra, dist, G.CONS_BOUNDS, 0.1)
        assert abs(float(val) - gold["value"]) <= 2e-6 * abs(gold["value"])
        for g, dg in zip(grads, gold["grads"]):
            assert not G.digest_close(dg, G.tensor_digest(g, n_head=12), 1e-5)
        return
    if kind == "TV":
        val, g = O.loss_tv(out, *G.TV_HW)
    elif kind.startswith("HDR"):
        val, g, _ = O.loss_hdr(out, gt, kc, float(opts["hdr_ff_sigma"]), float(opts["hdr_eps"]),
                               float(opts["hdr_ff_factor"]))
    elif kind == "LSL":
        val, g = O.loss_logspace(out, gt, float(opts["hdr_eps"]))
    else:
        val, g = O.LOSS_TRAIN[kind](out, gt)
    assert abs(float(val) - gold["value"]) <= 5e-6 * abs(gold["value"])
    assert not G.digest_close(gold["dout"], G.tensor_digest(g, n_head=12), 2e-5)
