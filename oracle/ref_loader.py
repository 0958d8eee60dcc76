"""Import the REAL reference modules from /root/reference (only present in the build container).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used to pin the restatement and to generate
tests/golden/*.npz; nothing on the GPU box may depend on it (`available()` is False there).
Recipe: SURVEY.md Appendix F — a synthetic package whose __path__ is the reference directory
(its own __init__ pulls ComfyUI), plus the single oracle patch for defect D1.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_DIR = os.environ.get("DRB_REFERENCE_DIR", "/root/reference")
_PKG = "dr_ref"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "CleanGeneralDIT.py"))


def load():
    """Returns (dit, cfg, mdl, pipe) reference modules with the head-flatten patch applied."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [REFERENCE_DIR]
        sys.modules[_PKG] = pkg
    dit = importlib.import_module(f"{_PKG}.CleanGeneralDIT")
    cfg = importlib.import_module(f"{_PKG}.diffusion_renderer_config")
    mdl = importlib.import_module(f"{_PKG}.model_diffusion_renderer")
    pipe = importlib.import_module(f"{_PKG}.diffusion_renderer_pipeline")
    if not getattr(dit.PytorchDotProductAttention, "_drb_patched", False):
        orig = dit.PytorchDotProductAttention.forward

        def patched(self, q, k, v, **kw):          # CleanGeneralDIT.py:199-203 returns 4-D; to_out needs 3-D
            return orig(self, q, k, v, **kw).flatten(2)

        dit.PytorchDotProductAttention.forward = patched
        dit.PytorchDotProductAttention._drb_patched = True
    return dit, cfg, mdl, pipe


def load_envmap():
    """The reference preprocess_envmap module with its absent third-party imports (nvdiffrast, imageio, cv2) stubbed:
    only its pure-torch functions are usable, which is all the oracle pins."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [REFERENCE_DIR]
        sys.modules[_PKG] = pkg
    for name in ("nvdiffrast", "nvdiffrast.torch", "imageio", "imageio.v3", "cv2", "OpenEXR", "Imath"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    return importlib.import_module(f"{_PKG}.preprocess_envmap")
