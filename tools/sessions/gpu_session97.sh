#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s97
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "shufflenet" 2>&1 | tail -3
timeout 900 python bench.py --model SlowFastShuffleNet --batch 256 --frames 16 --crop 112 --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_shufflenet.jsonl > $O/bench_shufflenet.json 2> $O/bench_shufflenet.err; python - $O/bench_shufflenet.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), {k:v['ms'] for k,v in list(d['kernel_breakdown'].items())[:7]})
PY
