#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s67
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python bench.py --model SlowFastMoibleNetV2 --batch 128 --steps 3 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_mobilenetv2.jsonl > $O/bench_mobilenetv2.json 2> $O/bench_mobilenetv2.err
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/s67/ops_mobilenetv2.jsonl')]
dw=[r for r in rows if r['kind'] in ('dwconv','conv_direct')]
dw.sort(key=lambda r:-r['ms'])
for r in dw[:30]: print(r['kind'], r['label'], round(r['ms'],3), "GB/s %.0f"%(r['bytes']/r['ms']/1e6), "GFLOP/s %.0f"%(r['flops']/r['ms']/1e6), r.get('why_direct',''))
PY
