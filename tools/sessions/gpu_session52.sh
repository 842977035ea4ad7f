#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s52
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s52/bench_b64.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_uint8_frames']['value'])
print(d['kernel_breakdown'])
PY
tail -3 $O/bench_b64.err
