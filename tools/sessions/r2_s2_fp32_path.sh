#!/bin/bash
# round 2: FP32-accurate path (cfg.ESF.PRECISION = "fp32"): kernel + model parity, then its throughput beside FP16
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s2
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fp32_path.py -q -s -m gpu > $O/pytest_fp32.log 2>&1; echo "pytest rc $?"
grep -E "rel err|max\|d\||passed|failed|Error|error" $O/pytest_fp32.log | head -60
timeout 600 python bench.py --precision fp32 --batch 16 --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline \
  --dump-ops $O/ops_fp32_b16.jsonl > $O/bench_fp32_b16.json 2> $O/bench_fp32_b16.err
tail -1 $O/bench_fp32_b16.json | cut -c1-400; tail -3 $O/bench_fp32_b16.err
