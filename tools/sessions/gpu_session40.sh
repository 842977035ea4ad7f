#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s40
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
for md in 0 1; do
ESF_ATTN_PINGPONG=$md timeout 900 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" 2>&1 | tail -1
ESF_ATTN_PINGPONG=$md timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_$md.jsonl > $O/bench_$md.json 2> $O/bench_$md.err
python - $O/ops_$md.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l)
    if r['kind']=='attention': print(r['label'], r['ms'])
PY
done
