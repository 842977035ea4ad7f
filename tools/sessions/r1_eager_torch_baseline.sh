#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s58
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for M in tf32 bf16; do
timeout 600 python tests/experiments/eager_torch_gpu.py 8 224 $M 2>&1 | tail -1 | tee -a $O/eager_torch_gpu.jsonl
done
timeout 600 python tests/experiments/eager_torch_gpu.py 16 224 tf32 slowfast_r50 2>&1 | tail -1 | tee -a $O/eager_torch_gpu.jsonl
