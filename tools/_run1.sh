mkdir -p gpurun_out
python tools/attn_ab.py 28160 32 2 > gpurun_out/attn_ab.log 2>&1 || { tail -5 gpurun_out/attn_ab.log; exit 1; }
tail -3 gpurun_out/attn_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 3 -c 1 -f -o gpurun_out/attn_final python tools/attn_ab.py 28160 32 2 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
