"""Phase breakdown of the attention kernel (tuning aid).  Needs a library built with -DDRB_ATTN_PROFILE:
    tools/build_prof.sh && DRB200_LIB=$PWD/tools/_prof/libdrb200_prof.so python tools/attn_profile.py
Prints, for CTA (0,0), the cycles softmax warp 4 and the MMA-issuing warp spent in each phase (see PROF(i) in attention.cu)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from drb200 import _lib, ops

S, H = 28160, 32
q = torch.randn(S, 3 * H * 128, device="cuda").bfloat16()
out = torch.empty(S, H * 128, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(q[:, :H * 128], q[:, H * 128:2 * H * 128], q[:, 2 * H * 128:], H, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q[:, :H * 128], q[:, H * 128:2 * H * 128], q[:, 2 * H * 128:], H, out=out)
e1.record()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 16)()
lib = _lib.load()
lib.drb_debug_attn_profile.argtypes = [ctypes.c_void_p]
assert lib.drb_debug_attn_profile(buf) == 0
v = list(buf)
n = S // 128
print(f"kernel {e0.elapsed_time(e1):.3f} ms; per kv-iteration cycles of CTA(0,0) ({n} iterations)")
names = ["loop/other", "wait s_full", "ldtm + rescale check + exp chunks 0,1 + sttm", "wait_ld + exp chunk 2", "-",
         "wait_st + fence + arrive half0", "exp chunk 3 + sttm", "sum + wait_st + fence + arrive half1"]
tot = sum(v[:8])
for nm, c in zip(names, v[:8]):
    if nm != "-":
        print(f"  softmax  {nm:40s} {c / n:8.1f}  ({100 * c / tot:4.1f}%)")
print(f"  softmax  total {tot / n:.1f} cycles/iteration")
names = ["issue + loop", "wait kv_full", "issue PV/QK (sum over t)", "wait p half0", "wait p half1", "-", "-", "-"]
tot = sum(v[8:16])
for nm, c in zip(names, v[8:16]):
    if nm != "-":
        print(f"  mma      {nm:40s} {c / n:8.1f}  ({100 * c / max(tot, 1):4.1f}%)")
print(f"  mma      total {tot / n:.1f} cycles/iteration")
