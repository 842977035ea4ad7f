#!/bin/bash
# round 2: tensor-core attention of the FP32-accurate plan (P and V as FP16 pairs) -- kernel + model parity, throughput
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s11
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fp32_path.py -q -s -m gpu > $O/pytest_fp32.log 2>&1; echo "pytest fp32 rc $?"
grep -E "rel err|passed|failed|Error|error|assert|timeout" $O/pytest_fp32.log | head -60
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" > $O/pytest_attn.log 2>&1; echo "pytest attention (16-bit) rc $?"; tail -2 $O/pytest_attn.log
timeout 600 python bench.py --precision fp32 --batch 16 --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline \
  --dump-ops $O/ops_fp32_b16.jsonl > $O/bench_fp32_b16.json 2> $O/bench_fp32_b16.err
tail -1 $O/bench_fp32_b16.json | cut -c1-300; tail -3 $O/bench_fp32_b16.err
python - <<PY
import json,collections
rows=[json.loads(l) for l in open("$O/ops_fp32_b16.jsonl")]
k=collections.Counter()
for r in rows: k[r["kind"]]+=r["ms"]
print({a:round(b,2) for a,b in k.most_common()})
PY
