#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s57
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_conv_igemm" 2>&1 | tail -25
for M in 0 1; do
ESF_IGEMM_THALO=$M timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_$M.jsonl > $O/bench_$M.json 2> $O/bench_$M.err
python - $O/ops_$M.jsonl $M <<'PY'
import json,sys
tot=0
for l in open(sys.argv[1]):
    r=json.loads(l)
    if r['kind']=='conv_igemm' and r['label'].startswith('3x1x1') and ('->16 ' in r['label'] or '->32 ' in r['label'] or '->64 ' in r['label']):
        tot+=r['ms']; print(sys.argv[2], r['label'], r['ms'])
print("halo mode", sys.argv[2], "fast 3x1x1 total", round(tot,3))
PY
done
