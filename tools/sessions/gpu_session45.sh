#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s45
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_model.py -x -q -m gpu > $O/gpu_tests.log 2>&1; tail -4 $O/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s45/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
PY
tail -2 $O/bench_b64.err
