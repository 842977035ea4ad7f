#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s13
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s13/build.log 2>&1
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/s13/ops_b64.jsonl > gpurun_out/s13/bench_b64.json 2> gpurun_out/s13/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s13/bench_b64.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d['roofline'])
for k,v in d['kernel_breakdown'].items(): print(k, v)
PY
tail -3 gpurun_out/s13/bench_b64.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 1 -c 1 -o gpurun_out/s13/prof_attn_v2_d32 python tools/prof_attn.py 32 8 56 3 tc 1 > gpurun_out/s13/ncu1.log 2>&1; tail -1 gpurun_out/s13/ncu1.log
