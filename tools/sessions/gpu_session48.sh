#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s48
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_frames.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s48/bench_b64.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_uint8_frames')})
PY
tail -3 $O/bench_b64.err
