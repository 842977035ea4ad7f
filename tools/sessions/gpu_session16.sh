#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s16
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv_igemm or conv_wfold or stem" > $O/kern.log 2>&1; tail -15 $O/kern.log
P="python tools/prof_conv.py"
$P 64 8 56 56 64 256 1 1 1 1 1 1 1
$P 64 8 56 56 64 256 1 1 1 1 1 1 0
$P 64 8 56 56 256 64 1 1 1 1 1 1 0
$P 64 8 28 28 128 512 1 1 1 1 1 1 1
$P 64 8 14 14 256 1024 1 1 1 1 1 1 1
$P 64 8 28 28 576 256 3 1 1 1 1 1 0
$P 64 8 56 56 64 64 1 3 3 1 1 1 0
$P 64 32 56 56 8 32 1 1 1 1 1 1 1
$P 64 32 28 28 16 64 1 1 1 1 1 1 1
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/model.log 2>&1; tail -8 $O/model.log
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s16/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(k, v)
PY
tail -3 $O/bench_b64.err
