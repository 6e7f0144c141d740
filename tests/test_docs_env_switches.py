"""Every environment switch the library or the Python host reads is listed in INTEGRATION.md (section 3b), so that a
maintainer binding the C ABI sees all of them in one place."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mri_implicit_neural_representations_b200")


def _switches():
    names = set()
    for base, pat in ((os.path.join(PKG, "csrc"), r'getenv\("(INR_[A-Z0-9_]+)"\)'),
                      (PKG, r'environ(?:\.get)?[\(\[]\s*"(INR_[A-Z0-9_]+)"')):
        for fn in os.listdir(base):
            path = os.path.join(base, fn)
            if os.path.isfile(path) and fn.endswith((".cu", ".cuh", ".py")):
                names.update(re.findall(pat, open(path).read()))
    return names


def test_every_environment_switch_is_documented():
    names = _switches()
    assert len(names) >= 15, names                      # the scan itself works
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
