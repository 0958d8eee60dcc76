#!/usr/bin/env python
"""bench.py — DiT denoise steps/s of the 7B inverse renderer on a synthetic 57x704x1280 clip (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload inverse7b|tiny]
                    [--parallel dp|cp] [--no-video] [--no-cpu-baseline]

A "step" is ONE EDM Euler denoise step of one G-buffer pass: the sigma-only AdaLN vectors, c_in scaling + patchify,
the 28-block GeneralDIT forward over S = 28 160 tokens and the unpatchify + Euler update (guidance 0, the node
default).  N > 1 (torchrun, one rank per GPU): the five G-buffer passes / independent clips are data-parallel, every
rank runs its own pass with replicated weights and no data-path collective, so scaling is "weak" and `value` is the
whole-job steps/s = N*K / max-over-ranks time.  `--parallel cp` instead splits ONE video's token sequence over the N GPUs
(context parallelism: Ulysses exchange fused into the kernels over NVLink peer memory, csrc/cp.cu): every rank works on
the same step, `value` = K / max time, scaling "strong".

Keys beyond the base contract:
  roofline     the dominant kernel (flash attention, 53 % of the forward's FLOPs): algorithmic FLOPs per launch
               (4*S*S*D) / its mean launch duration from CUDA events recorded inside the timed region, against the
               measured sustained bf16 peak in MEASURED_PEAKS.json.
  e2e          the same step through the public reference-facing call net.forward(x, timesteps, latent_condition,
               context_index) with pinned HOST buffers: H2D of x / condition / sigma / context index and D2H of F inside
               the timed region.
  cpu_baseline the oracle (CPU restatement of the reference) timed on the host cores on a bounded sample.
  step_tflops  whole-step achieved TFLOP/s (6.8132e14 algorithmic FLOP per step) and its fraction of the peak.
  video        seconds per inverse-rendered video measured through the pipeline API (CleanDiffusionRendererPipeline
               .generate_video x 5 G-buffer passes: H2D of the fp32 clip, tokenizer encode once, 15 Euler steps and one
               decode + uint8 post-process + D2H per pass); with dp over N GPUs each rank renders ceil(5/N) passes.
`--impl reference` times the reference's own CPU implementation of the path (the oracle port, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for everything else (NCCL prints its
# "NCCL version ..." banner to fd 1 from C), and the result line is written to a private duplicate of the real stdout.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


WORKLOADS = {
    # name: (model_channels, blocks, heads, (frames, height, width))
    "inverse7b": dict(D=4096, L=28, H=32, clip=(57, 704, 1280)),
    "tiny": dict(D=512, L=4, H=4, clip=(9, 256, 256)),           # BASELINE configs[0] (parity / CPU-runnable case)
    # BASELINE configs[4]: CV8x8x8 tokenizer encode + decode of a 121-frame 704x1280 clip per GPU (not the headline metric;
    # `python bench.py --workload tokenizer121` prints its own line: clips/s, TFLOP/s of the implicit-GEMM convolutions)
    "tokenizer121": dict(clip=(121, 704, 1280), enc_flop=3.566e13, dec_flop=6.128e13),
    "tokenizer57": dict(clip=(57, 704, 1280), enc_flop=1.764e13, dec_flop=3.014e13),
}
METRIC = "dit_denoise_steps_per_s"
UNIT = "steps/s"


def latent_shape(clip):
    f, h, w = clip
    return (16, (f - 1) // 8 + 1, h // 8, w // 8)


def flops_per_forward(D, L, S, c_in=33, as_written=False):
    """L*(8SD^2 + 4S^2D + 16SD^2) + 2SD*4C_in + 2SD*64; `as_written` adds the reference's two dead
    cross-attention GEMMs per block (2 * 2SD^2)."""
    per_block = 8 * S * D * D + 4 * S * S * D + 16 * S * D * D + (4 * S * D * D if as_written else 0)
    return L * per_block + 2 * S * D * 4 * c_in + 2 * S * D * 64


def peaks():
    p = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update(bf16_tflops=m["bf16_tflops"], bf16_tflops_sustained=m["bf16_tflops_sustained"], source="measured")
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(wl, n_timed: int, n_warm: int, tokens_div: int):
    """One FA-CA-MLP block of the reference algorithm (oracle port, bf16, all host threads) on S/tokens_div tokens;
    extrapolated to a full forward by the exact as-written FLOP ratio.  Returns (steps_per_s, info)."""
    import torch
    from oracle import dit_oracle as do
    from oracle.weights import DitDims, make_state_dict
    D, H = wl["D"], wl["H"]
    c, t, h, w = latent_shape(wl["clip"])
    S = t * (h // 2) * (w // 2)
    t_sub = max(1, t // tokens_div)
    S_sub = t_sub * (h // 2) * (w // 2)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = DitDims(model_channels=D, num_blocks=1, num_heads=H)
    sd = make_state_dict(dims, seed=0, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(S_sub, 1, D, generator=g).bfloat16()
    emb = torch.randn(1, D, generator=g).bfloat16()
    lora = (torch.randn(1, 3 * D, generator=g) * 0.1).bfloat16()
    ctx = torch.randn(1, 1, dims.crossattn_emb_channels, generator=g).bfloat16()
    ang = do.rope_angles(dims, t_sub, h // 2, w // 2, torch.bfloat16, "cpu", sd["net.pos_embedder.seq"])

    def one_block():
        hcur = x
        with torch.no_grad():
            for j, kind in enumerate(("fa", "ca", "mlp")):
                hcur = do.sub_block(sd, f"net.blocks.block0.blocks.{j}", kind, dims, hcur, emb, lora, ctx, ang)
        return hcur

    for _ in range(n_warm):
        one_block()
    times = []
    for _ in range(n_timed):
        t0 = time.perf_counter()
        one_block()
        times.append(time.perf_counter() - t0)
    t_block = sum(times) / len(times)
    f_block_sub = flops_per_forward(D, 1, S_sub, as_written=True) - 2 * S_sub * D * (4 * 33 + 64)
    f_forward = flops_per_forward(D, wl["L"], S, as_written=True)
    t_forward = t_block * f_forward / f_block_sub
    info = {"kind": "port", "cores": cores,
            "sample": f"1 FA-CA-MLP block (D={D}) of the oracle port in bf16 on {S_sub} of {S} tokens, {n_timed} timed runs of "
                      f"{t_block:.2f} s each, scaled to a {wl['L']}-block forward by the as-written FLOP ratio {f_forward / f_block_sub:.1f}",
            "cpu_tflops": f_block_sub / t_block / 1e12}
    return 1.0 / t_forward, info, sum(times)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    div = 1 if args.workload == "tiny" else 4
    value, info, total = cpu_reference_sample(wl, max(1, args.steps), max(1, args.warmup), div)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config_block(args, wl),
            "cpu_baseline": dict(info, value=value, unit=UNIT),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def config_block(args, wl):
    c, t, h, w = latent_shape(wl["clip"])
    return {"workload": f"{args.workload}: GeneralDIT D={wl['D']} L={wl['L']} heads={wl['H']}, inverse renderer, "
                        f"{wl['clip'][0]}x{wl['clip'][1]}x{wl['clip'][2]} clip -> latent 16x{t}x{h}x{w}, S={t * (h // 2) * (w // 2)} tokens, "
                        "guidance 0, 15-step sigma schedule, random-init weights",
            "parallelism": ("single GPU" if args.gpus <= 1 else
                            f"cp{args.gpus}: one video's tokens split over the GPUs, Ulysses exchange fused into kernels over NVLink peer memory"
                            if args.parallel == "cp" else
                            f"ring{args.gpus}: one video's tokens split over the GPUs, K/V blocks pulled from peer memory in ring order"
                            if args.parallel == "ring" else f"dp{args.gpus} over G-buffer passes (replicated weights, no collective)"),
            "l2": "working set per kernel (>= 230 MB activations + 32..134 MB weights) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ B200 arm
# DRAM traffic of one attention launch at S = 28 160, 32 heads, from the ncu --set full capture of this kernel
# (profiles/r01_attention_final_ncu_raw.txt: dram__bytes_read.sum 704.1 MB + dram__bytes_write.sum 239.6 MB; algorithmic 923 MB)
ATTN_NCU_TRAFFIC_BYTES = {(28160, 32): 943.6e6}


def random_tokenizer(torch, dev):
    """the CV8x8x8 tokenizer (105.6 M parameters) with random-init weights on `dev`"""
    from drb200.CleanVAE import AutoencoderKLCosmos, CleanVAE
    m = AutoencoderKLCosmos()
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    vae = CleanVAE(model=m)
    vae.to(dev)
    vae.reset_dtype(torch.bfloat16)
    return vae


def time_video(torch, dist, model, wl, rank, world, dev, cp_mode=False):
    """One inverse-rendered video through the pipeline API: dp = the five G-buffer passes dealt over the ranks; cp = every
    rank runs all five passes, each pass's token sequence split over the ranks (the tokenizer runs replicated).  Returns
    (seconds max over ranks, passes of this rank, h2d bytes, d2h bytes)."""
    from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
    f, hh, ww = wl["clip"]
    vae = random_tokenizer(torch, dev)
    pipe = CleanDiffusionRendererPipeline(checkpoint_dir="", checkpoint_name="", model_type="inverse", vae_instance=vae,
                                          model_instance=model, guidance=0.0, num_steps=15, seed=42)
    clip = (torch.rand(1, 3, f, hh, ww, generator=torch.Generator().manual_seed(1234)) * 2 - 1).pin_memory()   # host, fp32
    mine = list(range(5)) if cp_mode else [p for p in range(5) if p % world == rank]

    def render(passes, steps):
        pipe.num_steps = steps
        out = None
        with pipe.shared_conditions():
            for p in passes:
                batch = {"rgb": clip, "video": clip, "context_index": torch.full((1, 1), p, dtype=torch.long)}
                out = pipe.generate_video(batch, normalize_normal=(p == 3), seed=42)       # uint8 (1,T,H,W,3) on the host
        return out

    render(mine[:1] or [0], 1)                      # warm-up: allocations, tensor maps, tokenizer weight packing
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = render(mine, 15)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d = clip.numel() * 4 if mine else 0
    d2h = (out.size if out is not None else 0) * len(mine)
    pipe.vae_instance = None
    model.vae = None
    return dt.item(), len(mine), h2d, d2h


def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from drb200 import _lib
    from drb200 import diffusion_renderer_config as cfgm
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel

    cp_mode = args.parallel in ("cp", "ring") and world > 1
    f, hh, ww = wl["clip"]
    cfg = cfgm.get_inverse_renderer_config(hh, ww, f)
    cfg["model_type"] = "inverse"
    cfg["net"].update(model_channels=wl["D"], num_blocks=wl["L"], num_heads=wl["H"])
    with torch.device("meta"):
        model = CleanDiffusionRendererModel(cfg)
    model = model.to_empty(device=dev).to(torch.bfloat16)
    net = model.net.init_weights_(seed=0)
    c, t, h, w = latent_shape(wl["clip"])
    S = t * (h // 2) * (w // 2)
    # dp: rank r renders its own G-buffer pass of its own clip; cp: every rank holds the same clip and pass
    g = torch.Generator(device=dev).manual_seed(1234 + (0 if cp_mode else rank))
    cond = (torch.randn(1, 16, t, h, w, device=dev, generator=g) * 0.5).bfloat16()
    ctx_idx = torch.full((1, 1), 0 if cp_mode else rank % 5, dtype=torch.long, device=dev)
    model.scheduler.set_timesteps(15, device=dev)
    sig = model.scheduler.sigmas.contiguous()
    x0 = (torch.randn(1, 16, t, h, w, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()

    cp = None
    t0f, t1f = 0, t
    if cp_mode:
        from drb200.context_parallel import ContextParallel, shard_frames
        cp = ContextParallel(mode="ring" if args.parallel == "ring" else "ulysses")
        net.enable_context_parallel(cp)
        t0f, t1f = shard_frames(t, rank, world)
    tl = t1f - t0f
    net._ensure_packed()
    ws = net._workspace(tl, h, w, dev, cp)
    net.prepare_condition(ws, cond[:, :, t0f:t1f], tl, h, w)
    use_ca = net.prepare_context(ws, net.context_token(ctx_idx))
    x_start = x0[0][:, t0f:t1f].contiguous()
    x = x_start.clone()

    def step(i, timers=None):
        k = i % 15
        if k == 0:
            x.copy_(x_start)
        net.denoise_step(ws, x, sig[k:k + 1], sig[k + 1:k + 2], use_ca, timers)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    launches0 = _lib.LAUNCHES
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(args.steps):
            step(args.warmup + i, timers)
        e1.record()
        sync()
    launches = _lib.LAUNCHES - launches0
    ms = e0.elapsed_time(e1)
    attn_ms = statistics.mean(a.elapsed_time(b) for a, b in timers)
    if not torch.isfinite(x.float()).all():
        raise SystemExit("non-finite latent after the timed steps")

    # ---- e2e: the public net.forward call with pinned host buffers (H2D + D2H inside the timed region); under cp every
    # rank passes the same full tensors and receives the full F (frame slices are exchanged by the final all-gather)
    hx = x0.cpu().pin_memory()
    hcond = cond.cpu().pin_memory()
    hidx = ctx_idx.cpu().pin_memory()
    hsig = sig[:15].cpu().pin_memory()
    hout = torch.empty((1, 16, t, h, w), dtype=torch.bfloat16).pin_memory()

    def e2e_step(i):
        k = i % 15
        dx = hx.to(dev, non_blocking=True)
        dc = hcond.to(dev, non_blocking=True)
        di = hidx.to(dev, non_blocking=True)
        ds = hsig[k:k + 1].to(dev, non_blocking=True)
        out = net(x=model.scheduler.scale_model_input(dx, ds), timesteps=ds, latent_condition=dc, context_index=di)
        hout.copy_(out, non_blocking=True)

    n_e2e = max(2, min(args.steps, 5))
    e2e_step(0)
    sync()
    e0.record()
    for i in range(n_e2e):
        e2e_step(i)
    e1.record()
    sync()
    e2e_ms = e0.elapsed_time(e1)
    h2d = hx.numel() * 2 + hcond.numel() * 2 + hidx.numel() * 8 + 4
    d2h = hout.numel() * 2

    video = None
    if not args.no_video:
        v_s, v_passes, v_h2d, v_d2h = time_video(torch, dist, model, wl, rank, world, dev, cp_mode)
        video = {"s_per_video": v_s, "passes_per_rank": 5 if cp_mode else -(-5 // world), "passes_rank0": v_passes, "steps_per_pass": 15,
                 "h2d_bytes_rank0": v_h2d, "d2h_bytes_rank0": v_d2h,
                 "api": "CleanDiffusionRendererPipeline.generate_video per G-buffer pass (host fp32 clip in, host uint8 frames out), "
                        "tokenizer encode once + 15 Euler steps + decode + post-process per pass; random-init tokenizer"}

    t_loc = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = t_loc.tolist()
    if rank == 0:
        pk = peaks()
        units = 1 if cp_mode else world                       # videos-in-flight: cp works on one step together
        steps_per_s = units * args.steps / (ms_max / 1e3)
        F = flops_per_forward(wl["D"], wl["L"], S)
        heads_local = wl["H"] // world if cp_mode else wl["H"]
        # per layer and rank: ulysses = all S x S for H/P heads (one launch); ring = S/P rows x S keys for all heads (P launches,
        # timed together) — the same FLOPs
        attn_flops = 4.0 * S * S * 128 * heads_local
        peak = pk["bf16_tflops_sustained"]
        ach = attn_flops / (attn_ms / 1e3) / 1e12
        step_tf = F * steps_per_s / 1e12 / world              # per-GPU achieved TFLOP/s
        line = {
            "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if cp_mode else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_block(args, wl),
            "e2e": {"value": units * n_e2e / (e2e_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "CleanDiffusionRendererGeneralDIT.forward(x, timesteps, latent_condition, context_index) with pinned host tensors"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "attention_kernel (drb_attention_bf16)", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s", "frac": ach / peak,
                         "traffic": ATTN_NCU_TRAFFIC_BYTES.get((S, heads_local)),
                         "traffic_source": "profiles/r01_attention_final_ncu_raw.txt (ncu --set full, dram read + write per launch, bytes)",
                         "peak_source": pk["source"] + " sustained bf16",
                         "launch_ms": attn_ms, "launches_timed": len(timers), "flops_per_launch": attn_flops,
                         "share_of_step": attn_ms * wl["L"] / (ms / args.steps)},
            "step_tflops": {"achieved_per_gpu": step_tf, "flops_per_step": F, "frac_of_sustained_peak": step_tf / peak,
                            "frac_of_burst_peak": step_tf / pk["bf16_tflops"]},
            # one inverse video = 5 G-buffer passes x 15 steps (DiT only; the measured end-to-end figure is `video`)
            "s_per_video_dit_only": 5 * 15 / steps_per_s if cp_mode else -(-5 // world) * 15 / (steps_per_s / world),
            "clocks": clk.summary(),
        }
        if video is not None:
            line["video"] = video
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, info, _ = cpu_reference_sample(wl, 2, 1, 1 if args.workload == "tiny" else 4)
                line["cpu_baseline"] = dict(info, value=v, unit=UNIT)
            except Exception as e:   # the GPU number stands on its own; say why the CPU leg is missing
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    if cp is not None:
        cp.close()
    if world > 1:
        dist.destroy_process_group()


def run_tokenizer(args, wl):
    """encode + decode of one clip per GPU (data-parallel over clips, no collective): whole-job clips/s"""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from drb200 import _lib
    vae = random_tokenizer(torch, dev)
    f, hh, ww = wl["clip"]
    x = (torch.rand(1, 3, f, hh, ww, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)) * 2 - 1).bfloat16()

    def step():
        return vae.decode(vae.encode(x))

    for _ in range(max(3, args.warmup)):
        y = step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            y = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = peaks()
        per = ms.item() / args.steps
        tf = (wl["enc_flop"] + wl["dec_flop"]) / per / 1e9
        emit({
            "metric": "tokenizer_encode_decode_clips_per_s", "value": world * 1e3 / per, "unit": "clips/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": per, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: CV8x8x8 tokenizer encode + decode of a {f}x{hh}x{ww} clip per GPU, random-init weights",
                       "parallelism": f"dp{world} over clips (no collective)" if world > 1 else "single GPU",
                       "l2": "activations of 0.4 .. 1.8 GB per layer exceed the 126 MB L2; no explicit flush"},
            "gpu_launches": _lib.LAUNCHES - l0,
            "roofline": {"bound": "tensor", "kernel": "conv3d_kernel (drb_conv3d_cl), whole encode + decode", "achieved": tf,
                         "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"], "traffic": None,
                         "flops_per_step": wl["enc_flop"] + wl["dec_flop"]},
            "finite": bool(torch.isfinite(y.float()).all()), "clocks": clk.summary()})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="inverse7b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-video", action="store_true", help="skip the measured end-to-end video leg")
    ap.add_argument("--parallel", default="dp", choices=["dp", "cp", "ring"],
                    help="N > 1: dp = one G-buffer pass per GPU (weak scaling, default); cp = one video split over the GPUs with "
                         "the Ulysses head exchange fused into the kernels; ring = the same split with the ring K/V schedule")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.workload.startswith("tokenizer"):
        if args.impl == "reference":
            raise SystemExit("the CPU reference arm is defined for the DiT workloads only")
        run_tokenizer(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args, wl)


if __name__ == "__main__":
    main()
