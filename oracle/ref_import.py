"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference model in this container.

`/root/reference` is mounted read-only in the build container only; it does not exist on
the GPU box.  Nothing under `-m gpu` tests, `smoke()` or `bench.py` may call this at run
time -- it is used (a) to validate `oracle/convnext_oracle.py` and (b) by
`oracle/make_golden.py` to generate the committed fixtures in `tests/golden/`.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("ACX_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torchlibrosa_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, "src", "audioset_convnext_inf", "pytorch", "convnext.py"))


_REF_MOD = None


def import_reference_convnext():
    """Return the reference module `audioset_convnext_inf.pytorch.convnext` (unmodified
    source, /root/reference/src/audioset_convnext_inf/pytorch/convnext.py).

    The repo root carries an import ALIAS of the same package name (audioset_convnext_inf/ -> the B200 package, so
    that the reference's scripts run unmodified); the real reference is therefore imported with its own `src` first on
    sys.path and every `audioset_convnext_inf*` entry of sys.modules set aside, and both are restored afterwards."""
    global _REF_MOD
    if _REF_MOD is not None:
        return _REF_MOD
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    src = os.path.join(REFERENCE_ROOT, "src")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k == "audioset_convnext_inf" or k.startswith("audioset_convnext_inf.")}
    sys.path.insert(0, src)
    try:
        importlib.invalidate_caches()
        mod = importlib.import_module("audioset_convnext_inf.pytorch.convnext")
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(src)), mod.__file__
    finally:
        sys.path.remove(src)
        for k in [k for k in sys.modules if k == "audioset_convnext_inf" or k.startswith("audioset_convnext_inf.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _REF_MOD = mod
    return mod


def build_reference_tiny():
    """The configuration every reference caller uses (convnext.py:499-505,
    evaluate_convnext_on_audioset.py:23-29)."""
    mod = import_reference_convnext()
    return mod.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0,
                             after_stem_dim=[252, 56], use_speed_perturb=False)
