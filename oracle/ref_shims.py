"""Import the UNMODIFIED reference modules from /root/reference/src  --  TEST INFRASTRUCTURE ONLY.

The reference needs ``fastmri``, ``h5py``, ``matplotlib`` and ``skimage`` at import time; none is
installed here.  This module plants minimal ``sys.modules`` stand-ins (restated from the library
definitions: fastmri==0.3.0 fft2c/ifft2c/complex_abs/rss/to_tensor, skimage structural_similarity)
so that ``models.networks``, ``models.mfn``, ``models.wire2d``, ``models.regularization`` and
``metrics.losses`` import and run exactly as shipped.  Nothing here is reachable from the product
package, and nothing in the ``-m gpu`` tests / ``smoke()`` / ``bench.py`` needs it at run time
(the reference tree does not exist on the GPU box).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_SRC = os.environ.get("INR_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "models"))


def _plant_shims():
    import torch
    from . import inr_oracle as O

    if "fastmri" not in sys.modules:
        fm = types.ModuleType("fastmri")
        fm.fft2c, fm.ifft2c = O.fft2c, O.ifft2c
        fm.complex_abs, fm.rss = O.complex_abs, O.rss
        data = types.ModuleType("fastmri.data")
        tr = types.ModuleType("fastmri.data.transforms")

        def to_tensor(arr):
            import numpy as np
            if np.iscomplexobj(arr):
                arr = np.stack((arr.real, arr.imag), axis=-1)
            return torch.from_numpy(arr)

        tr.to_tensor = to_tensor
        data.transforms = tr
        fm.data = data
        sys.modules.update({"fastmri": fm, "fastmri.data": data, "fastmri.data.transforms": tr})
    for name in ("h5py", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        met = types.ModuleType("skimage.metrics")
        met.structural_similarity = lambda x, y, data_range=None: O.ssim(x, y)
        sk.metrics = met
        sys.modules.update({"skimage": sk, "skimage.metrics": met})


def load(*names):
    """Return the requested reference modules, e.g. load('models.networks', 'metrics.losses').

    Modules are loaded by file path under the private prefix ``inr_reference.`` so they can
    coexist in one process with this repo's own drop-in ``src/models/...`` tree (which uses the
    same top-level names on purpose)."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_SRC)
    _plant_shims()
    mods = []
    for n in names:
        alias = "inr_reference." + n
        if alias not in sys.modules:
            path = os.path.join(REF_SRC, *n.split(".")) + ".py"
            spec = importlib.util.spec_from_file_location(alias, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            spec.loader.exec_module(mod)
        mods.append(sys.modules[alias])
    return mods[0] if len(mods) == 1 else mods
