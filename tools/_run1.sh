mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tokenizer_gpu.py -x -q > gpurun_out/r02_tok_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_tok_tests.log
for pf in 1 0; do echo "== DRB_CONV_PREFETCH=$pf"; DRB_CONV_PREFETCH=$pf timeout 120 python tools/conv_t_probe.py 256; DRB_CONV_PREFETCH=$pf timeout 120 python tools/conv_t_probe.py 512 8 88 160; done > gpurun_out/r02_conv_t_probe_prefetch.log 2>&1; cat gpurun_out/r02_conv_t_probe_prefetch.log
timeout 300 python tools/vae_kernel_probe.py > gpurun_out/r02_vae_kernel_probe_g.log 2>&1; head -30 gpurun_out/r02_vae_kernel_probe_g.log
