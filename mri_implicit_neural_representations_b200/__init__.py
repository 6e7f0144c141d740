"""B200-native fitting engine for coordinate MLPs on multi-coil MRI data.

Python host side above the C ABI (include/inr_b200.h -> libinr_b200.so, hand-written sm_100a kernels).
There is no fallback path: touching any engine symbol without the built CUDA library raises InrError.
(`build` is importable on its own so the library can be (re)built before it is loaded.)"""

_ENGINE_SYMBOLS = {"InrError": "_lib", "lib": "_lib", "LIB_PATH": "_lib",
                   "ChainEngine": "engine", "Plan": "engine", "selftest_umma": "engine", "set_sm_budget": "engine"}


def __getattr__(name):
    mod = _ENGINE_SYMBOLS.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib
    return getattr(importlib.import_module("." + mod, __name__), name)
