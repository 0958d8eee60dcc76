mkdir -p gpurun_out
for k in 1 2; do
echo "--- old (cvt rounding, runtime modulus)"; DRB200_LIB=/root/repo/lib_old.so python tools/gn_probe.py | head -2
echo "--- new (ALU rounding, pow2 mask)"; python tools/gn_probe.py | head -2
done
echo "--- tokenizer old"; DRB200_LIB=/root/repo/lib_old.so python bench.py --workload tokenizer121 --steps 10 --warmup 3 --no-gpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])"
echo "--- tokenizer new"; python bench.py --workload tokenizer121 --steps 10 --warmup 3 --no-gpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])"
timeout 300 python -m pytest tests/test_tokenizer_gpu.py -x -q 2>&1 | tail -2
