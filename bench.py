#!/usr/bin/env python
"""bench.py — DiT denoise steps/s of the 7B inverse renderer on a synthetic 57x704x1280 clip (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload inverse7b|forward7b|tiny|tokenizer121]
                    [--parallel auto|cp|dp|ring] [--cp-batch B] [--no-video] [--no-cpu-baseline] [--no-gpu-baseline]

A "step" is ONE EDM Euler denoise step of one G-buffer pass: the sigma-only AdaLN vectors, c_in scaling + patchify,
the 28-block GeneralDIT forward over S = 28 160 tokens and the unpatchify + Euler update (guidance 0, the node default).

N = 1: one pass per step on one GPU.
N > 1 (torchrun, one rank per GPU), default `--parallel auto` = context parallelism: ONE video is split over the N GPUs —
every rank owns 1/N of the token sequence, the Ulysses head exchange around self-attention is fused into the QKV GEMM /
attention epilogues as P2P stores over NVLink (csrc/cp.cu) — and the five G-buffer passes of that video run batched along
the token-row axis (B = 5 sequences per transformer pass).  One timed iteration is therefore 5 steps; `value` =
5*K / max-over-ranks time, `ms_per_step` = iteration time / 5, scaling "strong" (the same video, N times the GPUs).  Before
anything is timed every rank checks, on a small net, that the context-parallel forward and the batched sampler are
bit-identical to the same model on one GPU (`cp_gate`).  The data-parallel alternative (rank r renders its own pass / clip,
replicated weights, no collective, weak scaling) is timed afterwards in the same job and reported under `dp`.

Keys beyond the base contract:
  roofline     the dominant kernel (flash attention, 53 % of the forward's FLOPs): algorithmic FLOPs per launch
               (4*S*S*128 per head and sequence) / its mean launch duration from CUDA events recorded inside the timed
               region, against the measured sustained bf16 peak in MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch
               from the ncu --set full capture recorded in profiles/attention_traffic.json, valid only while the sha256
               of csrc/attention.cu matches the one recorded there (else null + traffic_stale).
  e2e          the same metric through the public API with pinned HOST buffers, H2D of the inputs and D2H of the result
               inside the timed region: N = 1 net.forward(x, timesteps, latent_condition, context_index); N > 1
               model.sample_latent(...) of one step for the five batched passes.
  cpu_baseline the oracle (CPU restatement of the reference, pinned bit-exactly against it) timed on the host cores on a
               bounded sample: one FA-CA-MLP block at the full S = 28 160, x28 by the exact FLOP ratio (`extrapolated`).
  gpu_baseline the "existing Blackwell path": the reference's operator sequence through stock PyTorch (cuBLASLt, SDPA,
               unfused elementwise) for one block on the same GPU, run back to back for >= 3 s, x28 (N = 1 only).
  step_tflops  whole-step achieved TFLOP/s (6.8132e14 algorithmic FLOP per step) and its fraction of the peak.
  host_enqueue_ms_per_iteration   host time to enqueue one iteration's launches into an empty stream (ctypes calls + tensor-map
               encoding; no CUDA graph) — compare with ms_per_iteration: the host runs far ahead of the GPU.
  video        seconds per inverse-rendered video measured through the pipeline API (host fp32 clip in, host uint8
               frames out): N = 1 generate_video x 5 passes; N > 1 generate_video_passes (one batched sampler run).
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
    "inverse7b": dict(D=4096, L=28, H=32, clip=(57, 704, 1280), kind="inverse"),
    # BASELINE configs[2]: the forward renderer — 8 conditions (5 G-buffers + 3 environment maps) -> 136 condition
    # channels, K = 612 -> 616 patch GEMM, no context embedding (the cross-attention sub-block is the identity)
    "forward7b": dict(D=4096, L=28, H=32, clip=(57, 704, 1280), kind="forward"),
    "tiny": dict(D=512, L=4, H=4, clip=(9, 256, 256), kind="inverse"),   # BASELINE configs[0] (parity / CPU-runnable case)
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
def cpu_reference_sample(wl, n_timed: int, n_warm: int, budget_s: float = 150.0):
    """One FA-CA-MLP block of the reference algorithm (oracle port — pinned bit-exactly against the imported reference,
    tests/test_oracle_vs_reference.py — bf16, all host threads), scaled to a full forward by the exact as-written FLOP
    ratio.  The block runs at the full S when (n_timed + n_warm) blocks fit `budget_s`, else on S/2 or S/4 tokens (whole
    latent frames; attention is quadratic in S, so the ratio is computed for the sampled S).  Returns
    (steps_per_s, info, seconds per timed sample)."""
    import torch
    from oracle import dit_oracle as do
    from oracle.weights import DitDims, make_state_dict
    D, H = wl["D"], wl["H"]
    c, t, h, w = latent_shape(wl["clip"])
    S = t * (h // 2) * (w // 2)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = DitDims(model_channels=D, num_blocks=1, num_heads=H)
    sd = make_state_dict(dims, seed=0, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(1, D, generator=g).bfloat16()
    lora = (torch.randn(1, 3 * D, generator=g) * 0.1).bfloat16()
    ctx = torch.randn(1, 1, dims.crossattn_emb_channels, generator=g).bfloat16()

    def make(t_sub):
        S_sub = t_sub * (h // 2) * (w // 2)
        x = torch.randn(S_sub, 1, D, generator=g).bfloat16()
        ang = do.rope_angles(dims, t_sub, h // 2, w // 2, torch.bfloat16, "cpu", sd["net.pos_embedder.seq"])

        def one_block():
            hcur = x
            with torch.no_grad():
                for jj, kind in enumerate(("fa", "ca", "mlp")):
                    hcur = do.sub_block(sd, f"net.blocks.block0.blocks.{jj}", kind, dims, hcur, emb, lora, ctx, ang)
            return hcur
        return S_sub, one_block

    def block_flops(S_sub):
        return flops_per_forward(D, 1, S_sub, as_written=True) - 2 * S_sub * D * (4 * 33 + 64)

    # calibrate on a quarter of the frames, then take the largest sample that fits the budget
    t_q = max(1, t // 4)
    S_q, blk = make(t_q)
    blk()
    t0 = time.perf_counter()
    blk()
    t_cal = time.perf_counter() - t0
    t_sub = t_q
    for cand in (t, max(1, t // 2)):
        S_c = cand * (h // 2) * (w // 2)
        if t_cal * block_flops(S_c) / block_flops(S_q) * (n_timed + n_warm) <= budget_s:
            t_sub = cand
            break
    if t_sub != t_q:
        S_sub, blk = make(t_sub)
    else:
        S_sub = S_q
    for _ in range(n_warm):
        blk()
    times = []
    for _ in range(n_timed):
        t0 = time.perf_counter()
        blk()
        times.append(time.perf_counter() - t0)
    t_block = sum(times) / len(times)
    f_block_sub = block_flops(S_sub)
    f_forward = flops_per_forward(D, wl["L"], S, as_written=True)
    t_forward = t_block * f_forward / f_block_sub
    info = {"kind": "port", "cores": cores, "extrapolated": True,
            "port_pinned": "oracle port pinned bit-exactly (fp32, CPU) against the imported reference modules",
            "sample": f"1 FA-CA-MLP block (D={D}) of the oracle port in bf16 on {S_sub} of {S} tokens, {n_timed} timed runs of "
                      f"{t_block:.2f} s each, scaled to a {wl['L']}-block forward by the as-written FLOP ratio {f_forward / f_block_sub:.1f}",
            "sample_tokens": S_sub, "full_tokens": S, "s_per_sample": t_block, "flop_ratio": f_forward / f_block_sub,
            "cpu_tflops": f_block_sub / t_block / 1e12}
    return 1.0 / t_forward, info, t_block


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, info, t_sample = cpu_reference_sample(wl, max(1, args.steps), max(1, args.warmup))
    # `ms_per_step` is the measured wall time of one timed sample (so steps * ms_per_step is what this run really spent);
    # `value` is the extrapolated whole-forward rate the sample implies, in the B200 arm's unit
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * t_sample, "ms_per_full_step_extrapolated": 1000.0 / value,
            "extrapolated": True, "higher_is_better": True,
            # the same job as the B200 arm at this N (one video, total work fixed, when that arm runs context-parallel)
            "scaling": "strong" if args.gpus > 1 and args.parallel in ("auto", "cp", "ring") else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(config_block(args, wl, 1), parallelism="reference algorithm on the host cores of rank 0 (no GPU); the B200 arm "
                                                                  f"at --gpus {args.gpus} runs the same workload on {args.gpus} GPU(s)"),
            "cpu_baseline": dict(info, value=value, unit=UNIT),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def cond_channels(wl):
    return 16 if wl.get("kind", "inverse") == "inverse" else 136


def config_block(args, wl, world, mode="single", batch=1):
    c, t, h, w = latent_shape(wl["clip"])
    kind = wl.get("kind", "inverse")
    par = {"single": "single GPU",
           "cp": f"cp{world}: one video's tokens split over the {world} GPUs (Ulysses head exchange fused into the QKV GEMM / attention "
                 f"epilogues over NVLink peer memory), its {batch} G-buffer passes batched along the token-row axis",
           "ring": f"ring{world}: one video's tokens split over the GPUs, K/V blocks pulled from peer memory in ring order",
           "dp": f"dp{world}: {world} independent passes / clips, one per GPU (replicated weights, no collective)"}[mode]
    return {"workload": f"{args.workload}: GeneralDIT D={wl['D']} L={wl['L']} heads={wl['H']}, {kind} renderer "
                        f"({cond_channels(wl)} condition channels), {wl['clip'][0]}x{wl['clip'][1]}x{wl['clip'][2]} clip -> latent "
                        f"16x{t}x{h}x{w}, S={t * (h // 2) * (w // 2)} tokens, guidance 0, 15-step sigma schedule, random-init weights",
            "parallelism": par,
            "l2": "working set per kernel (>= 230 MB activations + 32..134 MB weights) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ B200 arm
def attention_traffic(S, heads):
    """DRAM bytes per attention launch from the committed ncu --set full capture — only while csrc/attention.cu is the file
    that was profiled (sha256 recorded next to the number)."""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "attention_traffic.json")) as f:
            rec = json.load(f)
        with open(os.path.join(ROOT, "diffusionrenderer-comfyui_b200", "csrc", "attention.cu"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
    except Exception as e:
        return None, f"unavailable: {e}", None
    val = rec.get("traffic_bytes", {}).get(f"{S}x{heads}")
    if val is None and f"{S}x32" in rec.get("traffic_bytes", {}) and rec.get("traffic_bytes_per_head"):
        val = rec["traffic_bytes_per_head"] * heads      # every (sequence, head) pair moves its own Q, K, V, O once
    if sha != rec.get("attention_cu_sha256"):
        return None, rec.get("source"), True
    return val, rec.get("source"), False


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


def cp_gate(torch, dist, cp, dev):
    """Correctness gate of the multi-GPU path, run where the driver runs the bench: on a small net the context-parallel
    forward, the context-parallel sampler and the sampler with batched passes + batched CFG must be BIT-IDENTICAL to the same
    model on one GPU (every rank computes the one-GPU reference itself).  Returns a dict for the result line."""
    from drb200 import diffusion_renderer_config as cfgm
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    cfg = cfgm.get_inverse_renderer_config(64, 96, 9)
    cfg["model_type"] = "inverse"
    cfg["net"].update(model_channels=1024, num_blocks=3, num_heads=8)
    torch.manual_seed(0)
    model = CleanDiffusionRendererModel(cfg).to(dev).to(torch.bfloat16)
    model.net.init_weights_(seed=1)
    T, H, W = 8, 16, 24                                   # 768 tokens; 96 per rank at P = 8
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(1, 16, T, H, W, device=dev, generator=g).bfloat16()
    cond = (torch.randn(1, 16, T, H, W, device=dev, generator=g) * 0.5).bfloat16()
    ci = torch.full((1, 1), 2, dtype=torch.long, device=dev)
    sigma = torch.tensor(1.26, device=dev)
    conds = [{"latent_condition": cond, "context_index": torch.full((1, 1), k, dtype=torch.long, device=dev)} for k in (0, 3, 4)]
    unconds = [{"latent_condition": torch.zeros_like(cond), "context_index": torch.zeros_like(ci)} for _ in conds]
    with torch.no_grad():
        model.scheduler.set_timesteps(3, device=dev)
        xt = x * model.scheduler.sigmas[0]
        ref_f = model.net(x=x, timesteps=sigma, latent_condition=cond, context_index=ci)
        ref_seq = torch.cat([model.sample_latent(xt, c, None) for c in conds])              # one pass at a time, one GPU
        ref_cfg = torch.cat([model.sample_latent(xt, c, u, guidance=2.0) for c, u in zip(conds, unconds)])
        model.net.enable_context_parallel(cp)
        got_f = model.net(x=x, timesteps=sigma, latent_condition=cond, context_index=ci)
        got_seq = model.sample_latent(xt, conds, None)                                       # batched passes, P GPUs
        got_cfg = model.sample_latent(xt, conds, unconds, guidance=2.0)                      # batched passes x (cond, uncond)
        model.net.enable_context_parallel(None)
    torch.cuda.synchronize()
    res = {"forward_identical": bool(torch.equal(got_f, ref_f)), "batched_sampler_identical": bool(torch.equal(got_seq, ref_seq)),
           "batched_cfg_sampler_identical": bool(torch.equal(got_cfg, ref_cfg)), "finite": bool(torch.isfinite(got_cfg.float()).all())}
    flag = torch.tensor([1 if all(res.values()) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["all_ranks_ok"] = bool(flag.item() == 1)
    res["what"] = ("tiny GeneralDIT (D=1024, 3 blocks, 8 heads, 768 tokens): cp forward / 3-pass batched sampler / batched CFG sampler "
                   "vs the same model on one GPU, torch.equal on every rank")
    del model
    torch.cuda.empty_cache()
    return res


def time_video(torch, dist, model, wl, rank, world, dev, mode):
    """One inverse-rendered video through the pipeline API.  single: generate_video per pass (what the node's loop does);
    dp: the five passes dealt over the ranks; cp: generate_video_passes — every rank calls it, the five passes run as one
    batched sampler over the split token sequence, pass p is decoded by rank p mod P, frames land on rank 0.
    Returns (seconds max over ranks, passes of this rank, h2d bytes, d2h bytes)."""
    from drb200.diffusion_renderer_pipeline import CleanDiffusionRendererPipeline
    f, hh, ww = wl["clip"]
    vae = random_tokenizer(torch, dev)
    pipe = CleanDiffusionRendererPipeline(checkpoint_dir="", checkpoint_name="", model_type="inverse", vae_instance=vae,
                                          model_instance=model, guidance=0.0, num_steps=15, seed=42)
    pipe.pinned_output = True      # as the node classes do: frames land in the pipeline's page-locked staging buffers
    clip = (torch.rand(1, 3, f, hh, ww, generator=torch.Generator().manual_seed(1234)) * 2 - 1).pin_memory()   # host, fp32
    mine = list(range(5)) if mode in ("cp", "ring", "single") else [p for p in range(5) if p % world == rank]

    def render(passes, steps):
        pipe.num_steps = steps
        if mode == "cp":
            outs = pipe.generate_video_passes({"rgb": clip, "video": clip}, passes, normalize_normal=[p == 3 for p in passes], seed=42)
            return [o.size for o in outs if o is not None]
        outs = []
        with pipe.shared_conditions():
            for p in passes:
                batch = {"rgb": clip, "video": clip, "context_index": torch.full((1, 1), p, dtype=torch.long)}
                arr = pipe.generate_video(batch, normalize_normal=(p == 3), seed=42)          # uint8 (1,T,H,W,3) on the host
                outs.append(arr.size)       # (a view of the staging buffer: consumed before the next pass, as the node does)
        return outs

    render(mine if mode == "cp" else (mine[:1] or [0]), 1)   # warm-up: allocations, tensor maps, tokenizer weight packing
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = render(mine, 15)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d = clip.numel() * 4 if mine else 0
    d2h = sum(outs)
    pipe.vae_instance = None
    model.vae = None
    return dt.item(), len(mine), h2d, d2h


def gpu_baseline(torch, min_seconds=3.0):
    """stock-PyTorch block (tools/torch_gpu_baseline.py) back to back for >= min_seconds on this GPU"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("torch_gpu_baseline", os.path.join(ROOT, "tools", "torch_gpu_baseline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.measure(28160, min_seconds)


def build_model(torch, wl, dev):
    from drb200 import diffusion_renderer_config as cfgm
    from drb200.model_diffusion_renderer import CleanDiffusionRendererModel
    f, hh, ww = wl["clip"]
    kind = wl.get("kind", "inverse")
    cfg = cfgm.get_config_by_model_type(kind, hh, ww, f)
    cfg["model_type"] = kind
    cfg["net"].update(model_channels=wl["D"], num_blocks=wl["L"], num_heads=wl["H"])
    with torch.device("meta"):
        model = CleanDiffusionRendererModel(cfg)
    model = model.to_empty(device=dev).to(torch.bfloat16)
    model.net.init_weights_(seed=0)
    return model


def timed_steps(torch, dist, world, local, net, ws, xs, sig, use_ca, steps, warmup, timers):
    """`warmup` untimed + `steps` timed denoise iterations on the latents `xs` ([N,16,T,H,W], restored every 15 steps);
    returns (device ms of the timed region, host enqueue ms, launches, clock summary)."""
    from drb200 import _lib
    x_start = xs.clone()

    def step(i, tm=None):
        k = i % 15
        if k == 0:
            xs.copy_(x_start)
        net.denoise_step(ws, xs, sig[k:k + 1], sig[k + 1:k + 2], use_ca, tm)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    launches0 = _lib.LAUNCHES
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(steps):
            step(warmup + i, timers)
        e1.record()
        sync()
    launched = _lib.LAUNCHES - launches0
    # host cost of enqueueing ONE iteration into an empty stream (inside the timed region the host runs ahead until the
    # driver's launch queue is full and then blocks on the GPU, which says nothing about the enqueue cost itself)
    h0 = time.perf_counter()
    step(warmup + steps)
    host_ms = (time.perf_counter() - h0) * 1e3
    sync()
    if not torch.isfinite(xs.float()).all():
        raise SystemExit("non-finite latent after the timed steps")
    return e0.elapsed_time(e1), host_ms, launched, clk.summary()


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
    c, t, h, w = latent_shape(wl["clip"])
    S = t * (h // 2) * (w // 2)
    mode = args.parallel
    if mode == "auto":
        mode = "single" if world == 1 else ("cp" if wl["H"] % world == 0 and t % world == 0 else "dp")
    if world == 1:
        mode = "single"
    # passes batched per iteration: 5 (all G-buffer passes of the video) from 4 GPUs up, where M = S/N rows per pass would
    # leave the GEMM grids with a short last wave; 1 at N = 2 (full waves anyway, and alternating attention / GEMM phases
    # run ~3 % faster under the power cap than long homogeneous phases: profiles/r02_cp2_batch_and_sync_sweep.log)
    batch = (args.cp_batch or (5 if world >= 4 else 1)) if mode == "cp" else 1

    cp, gate = None, None
    if mode in ("cp", "ring"):
        from drb200.context_parallel import ContextParallel, shard_frames
        cp = ContextParallel(mode="ring" if mode == "ring" else "ulysses")
        if mode == "cp":
            gate = cp_gate(torch, dist, cp, dev)
            if not gate["all_ranks_ok"]:
                if rank == 0:
                    emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "cp_gate failed", "cp_gate": gate})
                raise SystemExit("context-parallel correctness gate failed")

    model = build_model(torch, wl, dev)
    net = model.net
    cc = cond_channels(wl)
    # cp: every rank holds the same clip and the same five passes; single / dp: rank r renders its own pass of its own clip
    g = torch.Generator(device=dev).manual_seed(1234 + (0 if cp is not None else rank))
    cond = (torch.randn(1, cc, t, h, w, device=dev, generator=g) * 0.5).bfloat16()
    model.scheduler.set_timesteps(15, device=dev)
    sig = model.scheduler.sigmas.contiguous()
    x0 = (torch.randn(1, 16, t, h, w, device=dev, generator=g).bfloat16() * sig[0]).bfloat16()
    inverse = wl.get("kind", "inverse") == "inverse"

    def setup(cp_obj, nseq, first_pass):
        """workspace + constants of `nseq` passes (context indices first_pass, first_pass+1, ...) for this rank's frames"""
        t0f, t1f = (0, t) if cp_obj is None else shard_frames(t, rank, world)
        tl = t1f - t0f
        net.enable_context_parallel(cp_obj)
        net._ensure_packed()
        ws = net._workspace(tl, h, w, dev, cp_obj, batch=nseq)
        use_ca = False
        for b in range(nseq):
            net.prepare_condition(ws, cond[:, :, t0f:t1f], tl, h, w, b)
            ctx = net.context_token(torch.full((1, 1), (first_pass + b) % 5, dtype=torch.long, device=dev)) if inverse else None
            use_ca = net.prepare_context(ws, ctx, b)
        xs = x0[:, :, t0f:t1f].expand(nseq, -1, -1, -1, -1).contiguous().clone()
        return ws, xs, use_ca

    ws, xs, use_ca = setup(cp, batch, 0 if cp is not None else rank)
    timers = []
    ms, host_ms, launches, clocks = timed_steps(torch, dist, world, local, net, ws, xs, sig, use_ca, args.steps, args.warmup, timers)
    if cp is not None:
        cp.check()      # no in-kernel wait for a peer timed out
    attn_ms = statistics.mean(a.elapsed_time(b) for a, b in timers)

    # ---- e2e: the public call with pinned host buffers (H2D + D2H inside the timed region)
    hx = x0.cpu().pin_memory()
    hcond = cond.cpu().pin_memory()
    hsig = sig[:15].cpu().pin_memory()
    n_e2e = max(2, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if mode == "cp":
        # one public sampler call per timed iteration: model.sample_latent over the `batch` passes for ONE Euler step
        hidx = [torch.full((1, 1), b % 5, dtype=torch.long).pin_memory() for b in range(batch)]
        hout = torch.empty((batch, 16, t, h, w), dtype=torch.bfloat16).pin_memory()
        e2e_api = ("CleanDiffusionRendererModel.sample_latent(xt, [5 condition dicts]) for one Euler step per call, pinned host "
                   "tensors in (latent, latent condition, context indices, sigma pair), 5 host latents out")
        units_per_call = batch

        def e2e_step(i):
            k = i % 14
            model.scheduler.sigmas = hsig[k:k + 2].to(dev, non_blocking=True)
            dx = hx.to(dev, non_blocking=True)
            dc = hcond.to(dev, non_blocking=True)
            conds = [{"latent_condition": dc, "context_index": hi.to(dev, non_blocking=True)} for hi in hidx]
            hout.copy_(model.sample_latent(dx, conds, None), non_blocking=True)
        h2d = hx.numel() * 2 + hcond.numel() * 2 + batch * 8 + 8
        d2h = hout.numel() * 2
    else:
        hidx = torch.full((1, 1), rank % 5, dtype=torch.long).pin_memory()
        hout = torch.empty((1, 16, t, h, w), dtype=torch.bfloat16).pin_memory()
        e2e_api = "CleanDiffusionRendererGeneralDIT.forward(x, timesteps, latent_condition, context_index) with pinned host tensors"
        units_per_call = 1

        def e2e_step(i):
            k = i % 15
            dx = hx.to(dev, non_blocking=True)
            dc = hcond.to(dev, non_blocking=True)
            di = hidx.to(dev, non_blocking=True)
            ds = hsig[k:k + 1].to(dev, non_blocking=True)
            out = net(x=model.scheduler.scale_model_input(dx, ds), timesteps=ds, latent_condition=dc, context_index=di)
            hout.copy_(out, non_blocking=True)
        h2d = hx.numel() * 2 + hcond.numel() * 2 + hidx.numel() * 8 + 4
        d2h = hout.numel() * 2

    e2e_step(0)
    sync()
    e0.record()
    for i in range(n_e2e):
        e2e_step(i)
    e1.record()
    sync()
    e2e_ms = e0.elapsed_time(e1)
    model.scheduler.set_timesteps(15, device=dev)

    video = None
    if not args.no_video and inverse:
        v_s, v_passes, v_h2d, v_d2h = time_video(torch, dist, model, wl, rank, world, dev, mode)
        video = {"s_per_video": v_s, "passes_rank0": v_passes, "steps_per_pass": 15, "h2d_bytes_rank0": v_h2d, "d2h_bytes_rank0": v_d2h,
                 "api": ("CleanDiffusionRendererPipeline.generate_video_passes: host fp32 clip in, tokenizer encode, ONE 15-step sampler run "
                         "over the 5 batched passes, decode + post-process of pass p on rank p mod N, host uint8 frames out on rank 0"
                         if mode == "cp" else
                         "CleanDiffusionRendererPipeline.generate_video per G-buffer pass (host fp32 clip in, host uint8 frames out), "
                         "tokenizer encode once + 15 Euler steps + decode + post-process per pass") +
                        "; frames into the pipeline's pinned staging buffers (pinned_output, as the node classes set it); random-init tokenizer"}

    # ---- the data-parallel alternative, timed in the same job (N > 1 only): one independent pass per GPU
    dp = None
    if mode == "cp" and not args.no_dp_leg:
        ws_dp, xs_dp, use_ca_dp = setup(None, 1, rank)
        k_dp = max(2, min(args.steps, 4))
        ms_dp, _, _, _ = timed_steps(torch, dist, world, local, net, ws_dp, xs_dp, sig, use_ca_dp, k_dp, 3, None)
        t_dp = torch.tensor([ms_dp], device=dev, dtype=torch.float64)
        dist.all_reduce(t_dp, op=dist.ReduceOp.MAX)
        dp_sps = world * k_dp / (t_dp.item() / 1e3)
        dp = {"value": dp_sps, "unit": UNIT, "scaling": "weak", "steps": k_dp, "ms_per_step": t_dp.item() / k_dp,
              "parallelism": f"dp{world}: {world} independent passes / clips, one per GPU (replicated weights, no collective)",
              "s_per_video_dit_only": -(-5 // world) * 15 * (t_dp.item() / k_dp) / 1e3}

    t_loc = torch.tensor([ms, e2e_ms, host_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
    ms_max, e2e_max, host_max = t_loc.tolist()
    if rank == 0:
        pk = peaks()
        # single: one pass per iteration; cp: `batch` passes of one video per iteration (all ranks together); dp / ring-unbatched:
        # one pass per rank (dp) or one pass for all (ring)
        units = {"single": 1, "cp": batch, "ring": 1, "dp": world}[mode]
        steps_per_s = units * args.steps / (ms_max / 1e3)
        F = flops_per_forward(wl["D"], wl["L"], S, c_in=17 + cc)
        seqs_heads = batch * wl["H"] // world if mode == "cp" else wl["H"]
        # per attention stage and rank: cp = S x S for batch*H/P (sequence, head) pairs in one launch; single / dp = S x S for
        # H heads; ring = the same FLOPs as P block launches timed together
        attn_flops = 4.0 * S * S * 128 * (seqs_heads if mode != "ring" else wl["H"] // world)
        peak = pk["bf16_tflops_sustained"]
        ach = attn_flops / (attn_ms / 1e3) / 1e12
        gpus_sharing = world if mode in ("cp", "ring", "dp") else 1
        step_tf = F * steps_per_s / 1e12 / gpus_sharing        # per-GPU achieved TFLOP/s
        traffic, traffic_src, stale = attention_traffic(S, seqs_heads)
        line = {
            "metric": METRIC, "value": steps_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps / (batch if mode == "cp" else 1), "higher_is_better": True,
            "scaling": "strong" if mode in ("cp", "ring") else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_block(args, wl, world, mode, batch),
            "e2e": {"value": (units if mode != "cp" else units_per_call) * n_e2e / (e2e_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": e2e_api},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "attention_kernel (drb_attention_bf16*)", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": pk["source"] + " sustained bf16",
                         "launch_ms": attn_ms, "launches_timed": len(timers), "flops_per_launch": attn_flops,
                         "share_of_step": attn_ms * wl["L"] / (ms / args.steps)},
            "step_tflops": {"achieved_per_gpu": step_tf, "flops_per_step": F, "frac_of_sustained_peak": step_tf / peak,
                            "frac_of_burst_peak": step_tf / pk["bf16_tflops"]},
            "host_enqueue_ms_per_iteration": host_max,
            # one inverse video = 5 G-buffer passes x 15 steps (DiT only; the measured end-to-end figure is `video`)
            "s_per_video_dit_only": 5 * 15 / steps_per_s if mode != "dp" else -(-5 // world) * 15 / (steps_per_s / world),
            "clocks": clocks,
        }
        if stale:
            line["roofline"]["traffic_stale"] = "csrc/attention.cu changed since the ncu capture recorded in profiles/attention_traffic.json"
        if mode == "cp":
            line["ms_per_iteration"] = ms_max / args.steps
            line["passes_per_iteration"] = batch
            line["step_equivalents_timed"] = args.steps * batch      # K timed iterations x `batch` passes each; ms_per_step is per step-equivalent
            line["cp_gate"] = gate
        if dp is not None:
            line["dp"] = dp
        if video is not None:
            line["video"] = video
        if world == 1 and not args.no_gpu_baseline and args.workload == "inverse7b":
            try:
                del ws, xs
                net._ws.clear()
                torch.cuda.empty_cache()
                line["gpu_baseline"] = gpu_baseline(torch)
            except Exception as e:
                line["gpu_baseline"] = {"value": None, "unit": UNIT, "error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, info, _ = cpu_reference_sample(wl, 2, 1, budget_s=60.0)
                line["cpu_baseline"] = dict(info, value=v, unit=UNIT)
            except Exception as e:   # the GPU number stands on its own; say why the CPU leg is missing
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        emit(line)
    if cp is not None:
        net.enable_context_parallel(None)
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
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch block on the same GPU (N = 1)")
    ap.add_argument("--no-dp-leg", action="store_true", help="N > 1: skip the extra data-parallel measurement")
    ap.add_argument("--cp-batch", type=int, default=0, help="G-buffer passes batched per context-parallel iteration (0 = auto)")
    ap.add_argument("--parallel", default="auto", choices=["auto", "dp", "cp", "ring"],
                    help="N > 1: auto = cp when N divides the heads and the latent frames; cp = one video split over the GPUs "
                         "(Ulysses exchange fused into the kernels, passes batched); dp = one independent pass per GPU (weak "
                         "scaling); ring = the cp split with the ring K/V schedule")
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
