#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s61
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv or stem or wfold or nonlocal" 2>&1 | tail -3
python tools/prof_conv.py 64 8 56 56 64 32 1 1 1 1 1 1 0 5 fp16 0 2>&1 | tail -1
python tools/prof_conv.py 64 32 56 56 64 16 3 1 1 1 1 1 0 5 fp16 0 2>&1 | tail -1
python tools/prof_conv.py 64 8 56 56 64 256 1 1 1 1 1 1 1 5 fp16 0 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s61/bench_b64.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_uint8_frames']['value'])
print({k:v['ms'] for k,v in d['kernel_breakdown'].items()})
PY
