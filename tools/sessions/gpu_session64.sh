#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s64
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stem" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "dual_r50 or slowfast_r50 or i3d_r50" 2>&1 | tail -2
for M in 0 1; do
ESF_STEM_THALO=$M timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_$M.jsonl > $O/bench_$M.json 2> $O/bench_$M.err
python - $O/ops_$M.jsonl $O/bench_$M.json $M <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l)
    if r['kind']=='stem_igemm': print("halo", sys.argv[3], r['label'], r['ms'])
d=json.load(open(sys.argv[2])); print("halo", sys.argv[3], d['value'], d['ms_per_step'])
PY
done
