#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s22
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; tail -8 $O/gpu_tests.log
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s22/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['roofline'])
for k,v in d['kernel_breakdown'].items(): print(k, v)
PY
tail -3 $O/bench_b64.err
