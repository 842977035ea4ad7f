#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s5
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s5/build.log 2>&1
GROUP_TIMEOUT=200 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py -k "tcgen05" 2>&1 | tail -40
bash tools/gpu_bringup.sh tests/test_gpu_model.py -k "golden or default or stress" 2>&1 | tail -30
echo "=== bench batch 64"
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --dump-ops gpurun_out/s5/ops_b64.jsonl > gpurun_out/s5/bench_b64.json 2> gpurun_out/s5/bench_b64.err; tail -c 3500 gpurun_out/s5/bench_b64.json; tail -3 gpurun_out/s5/bench_b64.err
