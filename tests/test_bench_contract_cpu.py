"""bench.py's output contract, checked on the CPU-runnable arm (`--impl reference`, tiny workload): stdout carries exactly one
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "dit_denoise_steps_per_s" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""


def test_attention_traffic_record_belongs_to_the_committed_kernel_source():
    """`roofline.traffic` is only reported while csrc/attention.cu is the file ncu profiled: the record in profiles/ carries
    the sha256 of that source, and a stale record (kernel edited, capture not repeated) must be caught here, not at round end."""
    import hashlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "profiles", "attention_traffic.json")) as f:
        rec = json.load(f)
    with open(os.path.join(root, "diffusionrenderer-comfyui_b200", "csrc", "attention.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    assert rec["attention_cu_sha256"] == sha, "attention.cu changed: repeat the ncu --set full capture (tools/attn_ab.py 28160 32 2)"
    sys.path.insert(0, root)
    import bench
    val, src, stale = bench.attention_traffic(28160, 32)
    assert not stale and val == rec["traffic_bytes"]["28160x32"] and os.path.exists(os.path.join(root, src.split(" ")[0]))
    # traffic within 5 % of the algorithmic bytes (Q, K, V read once, O written once): no wasted re-reads
    assert abs(val / rec["algorithmic_bytes"] - 1.0) < 0.05
