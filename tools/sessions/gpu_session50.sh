#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s50
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "nln" 2>&1 | grep -E "rel err|passed|failed|Error|error" | head -12
for C in slow_nln_r50 slow_r50 i3d_nln_r50; do
timeout 900 python bench.py --case $C --frames 8 --crop 224 --batch 64 --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_$C.jsonl > $O/bench_$C.json 2> $O/bench_$C.err; python - $C <<'PY'
import json,sys
c=sys.argv[1]
try:
    d=json.load(open('gpurun_out/s50/bench_%s.json'%c))
    print(c,{k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['value'], d.get('e2e_uint8_frames',{}).get('value'))
    print(c,d['kernel_breakdown'])
except Exception as e:
    print(c,"FAILED",e); print(open('gpurun_out/s50/bench_%s.err'%c).read()[-1500:])
PY
done
