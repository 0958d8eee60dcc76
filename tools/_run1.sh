mkdir -p gpurun_out
PROBE_ITERS=1 timeout 800 ncu --set full --clock-control none --import-source on -k regex:conv3d_kernel -c 8 -f -o gpurun_out/conv_split_256_v2 python tools/conv_t_probe.py 256 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
