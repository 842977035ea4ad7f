#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s53
mkdir -p $O
ESF_NVCC_EXTRA=-DESF_ATTN_TIMING python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > $O/build.log 2>&1
for D in 8 32; do
  timeout 300 python tools/prof_attn.py $D 8 56 2 tc 1 2>&1 | grep -E "warp|attn" | head -24
done
