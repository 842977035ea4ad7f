#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s47
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for P in 0 1 2 3 4; do
  for D in 8 32; do
    ESF_ATTN_POLY=$P timeout 300 python tools/prof_attn.py $D 8 56 8 tc 5 2>&1 | tail -1 | sed "s/^/poly=$P /"
  done
done
for P in 2 4; do
  ESF_ATTN_POLY=$P timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn or attention" 2>&1 | tail -2 | sed "s/^/poly=$P /"
done
