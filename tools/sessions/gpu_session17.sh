#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s17
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn or attention" > $O/kern.log 2>&1; tail -15 $O/kern.log
for d in 32 8; do timeout 120 python tools/prof_attn.py $d 8 56 3 tc 3; done
timeout 1200 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/model.log 2>&1; tail -8 $O/model.log
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s17/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['roofline'])
for k,v in d['kernel_breakdown'].items(): print(k, v)
PY
tail -3 $O/bench_b64.err
