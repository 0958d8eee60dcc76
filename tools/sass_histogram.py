"""Per-kernel SASS instruction histogram of libdrb200.so (cuobjdump -sass): the Blackwell-native evidence — tcgen05 MMA
(UTCHMMA / UTCHMMA.2CTA), TMA (UTMALDG / UTMASTG, .MULTICAST), tensor-memory loads / stores (LDTM / STTM), tcgen05 commit
barriers (UTCBAR) — and the absence of legacy warp-level HMMA.  Runs without a GPU.

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusionrenderer-comfyui_b200", "libdrb200.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCBAR.MULTICAST", "SYNCS",
         "HMMA", "MUFU.EX2", "FMNMX", "FMNMX3", "FFMA2", "MEMBAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            if base == "UTMALDG" and "MULTICAST" in op:
                cur["UTMALDG.MULTICAST"] += 1
            if base == "UTCBAR" and "MULTICAST" in op:
                cur["UTCBAR.MULTICAST"] += 1
            if op.startswith("MUFU.EX2"):
                cur["MUFU.EX2"] += 1
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sha256 {sha}...), sm_100a; instruction counts per kernel (static)")
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print(f"{'kernel':60s} {'total':>7s} " + " ".join(f"{w:>{max(9, len(w))}s}" for w in WATCH))
    tot = collections.Counter()
    for k, c in kernels.items():
        n = re.sub(r"\(anonymous namespace\)::|drb::", "", names[k])
        n = n[:n.rfind("(")] if "(" in n else n          # drop the parameter list, keep the template arguments
        n = n.replace("(unsigned int)", "").replace("(bool)", "").replace("(int)", "").replace("void ", "").replace("<unnamed>::", "")
        print(f"{n[:60]:60s} {c['total']:7d} " + " ".join(f"{c.get(w, 0):{max(9, len(w))}d}" for w in WATCH))
        tot.update(c)
    print(f"{'ALL KERNELS':60s} {tot['total']:7d} " + " ".join(f"{tot.get(w, 0):{max(9, len(w))}d}" for w in WATCH))
    if tot.get("HMMA", 0):
        print("# WARNING: legacy HMMA present", file=sys.stderr)


if __name__ == "__main__":
    main()
