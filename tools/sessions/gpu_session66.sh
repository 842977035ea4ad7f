#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s66
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "global_mean" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "ghostnet" 2>&1 | grep -E "rel err|passed|failed" | head
timeout 900 python bench.py --model SlowFastGhostNet --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_ghostnet.json 2> $O/bench_ghostnet.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s66/bench_ghostnet.json'))
print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
