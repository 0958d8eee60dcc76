"""`drb200` — importable alias of the product package `diffusionrenderer-comfyui_b200/` (a directory name Python
cannot import directly, but one ComfyUI can load as a custom-node folder).  Submodules resolve into that directory:
`import drb200.ops` loads `diffusionrenderer-comfyui_b200/ops.py`."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diffusionrenderer-comfyui_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
