"""First-contact diagnostics for the sm_100a kernels: runs each check in its own subprocess (a trap/illegal access in
one kernel must not poison the CUDA context of the others) and prints error statistics, not just pass/fail.

    python tools/gpu_probe.py [name ...]        # on a B200 (under gpurun); writes gpurun_out/probe.log
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _stats(name, got, ref):
    import torch
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    rel = (got - ref).norm() / ref.norm().clamp_min(1e-30)
    print(f"  {name}: rel_l2={rel.item():.3e} max_abs={d.max().item():.3e} ref_absmax={ref.abs().max().item():.3e} "
          f"nan={int(torch.isnan(got).sum())} exact={(got == ref).float().mean().item():.4f}", flush=True)
    return rel.item()


def check_gemm(cta_group, epi, M, N, K):
    import torch
    from drb200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    resid = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    acc = (a.float() @ w.float().t())
    if epi == 0:
        ref = acc.bfloat16()
        out = ops.gemm(a, w, cta_group=cta_group)
    elif epi == 1:
        ref = torch.nn.functional.gelu(acc.bfloat16())
        out = ops.gemm(a, w, epilogue=1, cta_group=cta_group)
    else:
        ref = resid + gate[None, :] * acc.bfloat16()
        out = ops.gemm(a, w, epilogue=2, resid=resid, gate=gate, cta_group=cta_group)
    torch.cuda.synchronize()
    r = _stats(f"gemm cg={cta_group} epi={epi} {M}x{N}x{K}", out, ref)
    if r > 1e-2:
        # localise: per 128x64 block error map
        d = (out.float() - ref.float()).abs()
        bm, bn = min(8, (M + 127) // 128), min(8, (N + 63) // 64)
        for i in range(bm):
            print("   ", " ".join(f"{d[i*128:(i+1)*128, j*64:(j+1)*64].max().item():8.2e}" for j in range(bn)))
    return r


def check_gemm_perf(cta_group, M, N, K, iters=20):
    import torch
    from drb200 import ops
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, out=out, cta_group=cta_group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, w, out=out, cta_group=cta_group)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  gemm perf cg={cta_group} {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    for _ in range(3):
        torch.matmul(a, w.t(), out=out)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  cuBLAS (torch.matmul) {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


def check_attention(S, H, Skv=None):
    import torch
    from drb200 import ops
    Skv = S if Skv is None else Skv
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(S, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    out = ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    qf = q.float().view(S, H, 128).transpose(0, 1)
    kf = k.float().view(Skv, H, 128).transpose(0, 1)
    vf = v.float().view(Skv, H, 128).transpose(0, 1)
    ref = torch.softmax(qf @ kf.transpose(1, 2) / 128 ** 0.5, dim=-1) @ vf
    ref = ref.transpose(0, 1).reshape(S, H * 128)
    r = _stats(f"attention S={S} Skv={Skv} H={H}", out, ref)
    if r > 2e-2:
        d = (out.float() - ref).abs()
        for i in range(min(4, (S + 127) // 128)):
            print("   rows", i * 128, " ".join(f"{d[i*128:(i+1)*128, j*64:(j+1)*64].max().item():8.2e}"
                                              for j in range(min(8, H * 2))))
        print("   out[0,:8]", out[0, :8].float().tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
    return r


def check_attention_perf(S, H, iters=5):
    import torch
    from drb200 import ops
    q = torch.randn(S, 3 * H * 128, device="cuda").bfloat16()
    qv, kv, vv = q[:, :H * 128], q[:, H * 128:2 * H * 128], q[:, 2 * H * 128:]
    out = torch.empty(S, H * 128, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention(qv, kv, vv, H, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attention(qv, kv, vv, H, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * S * S * 128 * H
    print(f"  attention perf S={S} H={H}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    import torch.nn.functional as F
    q4 = qv.reshape(S, H, 128).permute(1, 0, 2)[None]
    k4 = kv.reshape(S, H, 128).permute(1, 0, 2)[None]
    v4 = vv.reshape(S, H, 128).permute(1, 0, 2)[None]
    for _ in range(2):
        F.scaled_dot_product_attention(q4, k4, v4)
    e0.record()
    for _ in range(iters):
        F.scaled_dot_product_attention(q4, k4, v4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"  torch SDPA S={S} H={H}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


CHECKS = {
    "gemm1_small": lambda: check_gemm(1, 0, 256, 256, 128),
    "gemm1_k": lambda: check_gemm(1, 0, 128, 256, 1024),
    "gemm1_ragged": lambda: check_gemm(1, 0, 300, 520, 136),
    "gemm1_gelu": lambda: check_gemm(1, 1, 512, 512, 512),
    "gemm1_gated": lambda: check_gemm(1, 2, 512, 512, 512),
    "gemm1_big": lambda: check_gemm(1, 0, 4096, 4096, 4096),
    "gemm2_small": lambda: check_gemm(2, 0, 256, 256, 128),
    "gemm2_ragged": lambda: check_gemm(2, 0, 300, 520, 136),
    "gemm2_gated": lambda: check_gemm(2, 2, 512, 512, 512),
    "gemm2_big": lambda: check_gemm(2, 0, 4096, 4096, 4096),
    "attn_128": lambda: check_attention(128, 1),
    "attn_256": lambda: check_attention(256, 1),
    "attn_512": lambda: check_attention(512, 4),
    "attn_ragged": lambda: check_attention(48, 2),
    "attn_1000": lambda: check_attention(1000, 2),
    "attn_4096": lambda: check_attention(4096, 2),
    "perf_gemm1": lambda: check_gemm_perf(1, 28160, 4096, 4096),
    "perf_gemm2": lambda: check_gemm_perf(2, 28160, 4096, 4096),
    "perf_gemm2_mlp1": lambda: check_gemm_perf(2, 28160, 16384, 4096, iters=8),
    "perf_gemm2_mlp2": lambda: check_gemm_perf(2, 28160, 4096, 16384, iters=8),
    "perf_attn": lambda: check_attention_perf(28160, 32),
}


def main():
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 1 and os.environ.get("DRB_PROBE_CHILD"):
        t0 = time.time()
        CHECKS[names[0]]()
        print(f"  [{names[0]} done in {time.time() - t0:.1f}s]", flush=True)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")
    env = dict(os.environ, DRB_PROBE_CHILD="1")
    for n in names:
        hdr = f"== {n}"
        print(hdr, flush=True)
        log.write(hdr + "\n")
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, capture_output=True, text=True,
                               timeout=180)
            out = r.stdout + ("" if r.returncode == 0 else f"  EXIT {r.returncode}\n" + r.stderr[-1500:])
        except subprocess.TimeoutExpired as e:
            out = f"  TIMEOUT\n{(e.stdout or b'')[-500:]}\n"
        print(out, flush=True)
        log.write(out + "\n")
        log.flush()


if __name__ == "__main__":
    main()
