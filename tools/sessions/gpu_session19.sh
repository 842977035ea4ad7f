#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s19
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for v in 3 9 11 17 19 25 27; do
  echo "variant $v"
  ESF_ATTN_VARIANT=$v timeout 120 python tools/prof_attn.py 32 8 56 3 tc 3
  ESF_ATTN_VARIANT=$v timeout 120 python tools/prof_attn.py 8 8 56 3 tc 3
done
for v in 19 27; do
ESF_ATTN_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" 2>&1 | tail -3
done
