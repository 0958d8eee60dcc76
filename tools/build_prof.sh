#!/bin/bash
# tuning build of libdrb200 with in-kernel phase counters (never shipped / never loaded by default)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_prof
SRC=diffusionrenderer-comfyui_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDRB_ATTN_PROFILE ${DRB_NVCC_EXTRA} -shared \
  -o tools/_prof/libdrb200_prof.so $SRC/*.cu
echo tools/_prof/libdrb200_prof.so
