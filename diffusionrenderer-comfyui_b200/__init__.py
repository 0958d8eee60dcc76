"""B200-native (sm_100a) implementation of the DiffusionRenderer denoising hot path, behind the reference's own
Python surface (reference `__init__.py:1-3` re-exports the ComfyUI node registry).

Everything heavy is imported lazily so that `import drb200` works on a box without a GPU (the CPU test tier only
needs the host logic and the symbol table of libdrb200.so)."""

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]


def __getattr__(name):
    if name in __all__:
        from . import nodes
        return getattr(nodes, name)
    raise AttributeError(name)
