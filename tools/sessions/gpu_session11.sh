#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s11
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s11/build.log 2>&1
echo "=== full gpu suite (as the driver runs it)"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s11/pytest_gpu.log 2>&1; echo rc=$?; tail -5 gpurun_out/s11/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench fp16 batch 64"
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --dump-ops gpurun_out/s11/ops_b64.jsonl > gpurun_out/s11/bench_b64.json 2> gpurun_out/s11/bench_b64.err; tail -c 1800 gpurun_out/s11/bench_b64.json; tail -3 gpurun_out/s11/bench_b64.err
