#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s36
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wfold or attention" > $O/t.log 2>&1; tail -3 $O/t.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "golden" > $O/t2.log 2>&1; tail -3 $O/t2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s36/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k,v in d['kernel_breakdown'].items(): print(k, v)
PY
tail -2 $O/bench_b64.err
