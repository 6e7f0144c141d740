"""Drop-in loss classes (src/metrics/losses.py) against values generated from the UNMODIFIED reference classes
(oracle/make_golden_dropin_losses.py -> tests/golden/dropin_losses.json): this pin also holds where the reference tree is
absent (the live comparison is tests/test_losses_vs_reference.py)."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dropin_losses_match_reference_golden():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_golden_dropin_losses as MG
    spec = importlib.util.spec_from_file_location("inr_src_losses_g", os.path.join(ROOT, "src", "metrics", "losses.py"))
    ours = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ours)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "dropin_losses.json")))
    got = MG.evaluate(ours)
    assert sorted(got) == sorted(gold)
    for name, g in gold.items():
        for key, want in g.items():
            have = got[name][key]
            if isinstance(want, list):
                for a, b in zip(have, want):
                    assert abs(a - b) <= 2e-5 * max(abs(b), 1e-6), (name, key, a, b)
            else:
                assert abs(have - want) <= 2e-5 * max(abs(want), 1e-9), (name, key, have, want)
