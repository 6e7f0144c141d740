"""Drop-in for the reference's src/models/networks.py: same class names, constructor arguments and state_dict
keys; forward/backward run on the B200 engine (mri_implicit_neural_representations_b200)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from mri_implicit_neural_representations_b200.modules import FFN, SIREN, WIRE, Positional_Encoder  # noqa: E402,F401
