#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s41
mkdir -p $O
for pp in 0 1; do
echo "== pingpong $pp d=32"; ESF_ATTN_PINGPONG=$pp timeout 120 python tools/prof_attn.py 32 8 56 3 tc 1 2>&1 | tail -18
done
echo "== pingpong 0 d=8"; ESF_ATTN_PINGPONG=0 timeout 120 python tools/prof_attn.py 8 8 56 3 tc 1 2>&1 | tail -18
