"""Event timeline of ONE context-parallel iteration (7B, 57x704x1280, B batched passes) on every rank: CUDA events around every
C-ABI call, written as CSV (rank, index, call, start_ms, dur_ms) plus a per-kernel summary — the stand-in for an nsys timeline
(nsys is not installed in this image).  Shows what the two barrier kernels per block cost (launch + wait for the slowest rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/cp_timeline.py [--batch 5]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/r02_cp_iteration_timeline")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from drb200 import _lib
    from drb200 import diffusion_renderer_config as cfgm
    from drb200.context_parallel import ContextParallel
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    cp = ContextParallel()
    cfg = cfgm.get_inverse_renderer_config(704, 1280, 57)
    cfg["model_type"] = "inverse"
    with torch.device("meta"):
        model = CleanDiffusionRendererModel(cfg)
    model = model.to_empty(device=dev).to(torch.bfloat16)
    net = model.net.init_weights_(seed=0)
    net.enable_context_parallel(cp)
    net._ensure_packed()
    T, H, W, B = 8, 88, 160, args.batch
    Tl = T // world
    ws = net._workspace(Tl, H, W, dev, cp, batch=B)
    g = torch.Generator(device=dev).manual_seed(1234)
    cond = (torch.randn(1, 16, T, H, W, device=dev, generator=g) * 0.5).bfloat16()
    model.scheduler.set_timesteps(15, device=dev)
    sig = model.scheduler.sigmas.contiguous()
    x = (torch.randn(16, T, H, W, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()[:, rank * Tl:(rank + 1) * Tl]
    x = x.unsqueeze(0).expand(B, -1, -1, -1, -1).contiguous()
    for b in range(B):
        net.prepare_condition(ws, cond[:, :, rank * Tl:(rank + 1) * Tl], Tl, H, W, b)
        use_ca = net.prepare_context(ws, net.context_token(torch.full((1, 1), b % 5, dtype=torch.long, device=dev)), b)
    for i in range(3):
        net.denoise_step(ws, x, sig[i:i + 1], sig[i + 1:i + 2], use_ca)
    dist.barrier()
    torch.cuda.synchronize()
    real_call, recs = _lib.call, []

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(name, *a)
        e1.record()
        key = name.replace("drb_", "")
        if name == "drb_gemm_bf16":
            key = f"gemm N={a[7]} K={a[8]} epi={a[9]}"
        recs.append((key, e0, e1))

    _lib.call = timed_call
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    net.denoise_step(ws, x, sig[3:4], sig[4:5], use_ca)
    torch.cuda.synchronize()
    _lib.call = real_call
    rows = [(k, t0.elapsed_time(a), a.elapsed_time(b)) for k, a, b in recs]
    total = rows[-1][1] + rows[-1][2]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(f"{args.out}_rank{rank}.csv", "w") as f:
        f.write("rank,index,call,start_ms,dur_ms\n")
        for i, (k, s, d) in enumerate(rows):
            f.write(f"{rank},{i},{k},{s:.4f},{d:.4f}\n")
    agg = collections.OrderedDict()
    for k, s, d in rows:
        t, c = agg.get(k, (0.0, 0))
        agg[k] = (t + d, c + 1)
    busy = sum(t for t, _ in agg.values())
    lines = [f"rank {rank}/{world}: one iteration of {B} batched passes = {total:.2f} ms wall ({total / B:.2f} ms per step), "
             f"{busy:.2f} ms inside calls, {total - busy:.2f} ms between them (event overhead + launch gaps)"]
    for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"  {k:34s} x{c:4d} {t:9.3f} ms {100 * t / total:5.1f}%")
    gathered = [None] * world
    dist.all_gather_object(gathered, "\n".join(lines))
    if rank == 0:
        with open(f"{args.out}_summary.txt", "w") as f:
            f.write("# tools/cp_timeline.py: CUDA events around every C-ABI call of one context-parallel iteration (barrier kernels: launch +\n"
                    "# wait for the slowest rank)\n" + "\n".join(gathered) + "\n")
        print(gathered[0])
        print(gathered[-1])
    cp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
