#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s23
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "clip_stream or reload" > $O/gpu_tests.log 2>&1; tail -8 $O/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s23/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
print(d.get('clocks'))
PY
tail -3 $O/bench_b64.err
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
