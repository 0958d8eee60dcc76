"""The oracle is test infrastructure: nothing in the product package, tools/ or the B200 arm of bench.py may import it."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom):
            yield node.module or ""


def test_product_and_tools_never_import_the_oracle():
    bad = []
    for sub in ("diffusionrenderer-comfyui_b200", "tools", "drb200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith(".py"):
                    p = os.path.join(dirpath, f)
                    bad += [(p, m) for m in _imports(p) if m.split(".")[0] in ("oracle", "tests")]
    assert not bad, bad


def test_bench_only_touches_the_oracle_in_the_cpu_reference_leg():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            uses = [m for sub in ast.walk(node) for m in ([a.name for a in sub.names] if isinstance(sub, ast.Import) else
                                                          [sub.module or ""] if isinstance(sub, ast.ImportFrom) else [])
                    if m.split(".")[0] == "oracle"]
            assert not uses or node.name == "cpu_reference_sample", (node.name, uses)
    top = [m for node in tree.body for m in ([a.name for a in node.names] if isinstance(node, ast.Import) else
                                             [node.module or ""] if isinstance(node, ast.ImportFrom) else [])]
    assert not any(m.split(".")[0] == "oracle" for m in top)


def test_integration_guide_names_every_declared_entry_point():
    """INTEGRATION.md is the maintainer's map from reference call sites to the C ABI: no declared symbol may be missing"""
    import re
    with open(os.path.join(ROOT, "include", "drb200.h")) as f:
        declared = set(re.findall(r"\b(drb_[a-z0-9_]+)\s*\(", f.read()))
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        guide = f.read()
    missing = sorted(s for s in declared if s not in guide)
    assert len(declared) >= 48 and not missing, missing
