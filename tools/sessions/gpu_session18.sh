#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s18
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for v in 0 1 2 3 4 5 6 7; do
  echo "variant $v"
  ESF_ATTN_VARIANT=$v timeout 120 python tools/prof_attn.py 32 8 56 3 tc 3
  ESF_ATTN_VARIANT=$v timeout 120 python tools/prof_attn.py 8 8 56 3 tc 3
done
ESF_ATTN_VARIANT=0 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" 2>&1 | tail -3
ESF_ATTN_VARIANT=4 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" 2>&1 | tail -3
