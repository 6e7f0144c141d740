"""Drop-in for the reference's src/models/wire2d.py: WIRE2D runs on the B200 engine (same constructor argument and
state_dict keys)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mri_implicit_neural_representations_b200.modules import WIRE2D  # noqa: E402,F401
