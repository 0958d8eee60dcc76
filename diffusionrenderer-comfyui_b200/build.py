"""In-tree build of libdrb200.so (the C-ABI library declared in include/drb200.h) for sm_100a.

`python -m drb200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the resulting
`libdrb200.so` sits next to this file (git-ignored) so that it travels with the tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdrb200.so")
OBJ_DIR = os.path.join(HERE, "build")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdrb200.so cannot be built (no CPU fallback exists)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths) -> str:
    """content hash of the sources, keyed by their path RELATIVE to the repository (the stamp must survive the tree being
    copied to another directory — the GPU box receives a snapshot under a scratch path)"""
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    for p in sorted(paths, key=lambda q: os.path.relpath(q, root)):
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, root).encode() + b"\0" + f.read())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libdrb200.so.  Returns the library path."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "drb200.h"))
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    try:
        nvcc = _nvcc()
    except RuntimeError:
        if os.path.exists(LIB) and not force:   # a box without the toolkit: use the library that travelled with the tree
            sys.stderr.write(f"build.py: nvcc not found and the source stamp does not match; reusing the existing {LIB}\n")
            return LIB
        raise
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nvcc, *ARCH_FLAGS, "-shared", "-o", LIB, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
