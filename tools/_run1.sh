mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tokenizer_gpu.py tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -3
python tools/vae_kernel_probe.py > gpurun_out/probe_final.log 2>&1; head -1 gpurun_out/probe_final.log
python bench.py --workload tokenizer121 --steps 10 --warmup 3 --no-gpu-baseline 2>/dev/null | tail -1 > gpurun_out/tok121.json; python -c "import json; d=json.loads(open('gpurun_out/tok121.json').read()); print('tokenizer121', d['ms_per_step'], d['roofline'])"
python bench.py --workload tokenizer57 --steps 10 --warmup 3 --no-gpu-baseline 2>/dev/null | tail -1 > gpurun_out/tok57.json; python -c "import json; d=json.loads(open('gpurun_out/tok57.json').read()); print('tokenizer57', d['ms_per_step'], d['roofline'])"
