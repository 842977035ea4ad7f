#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s7
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s7/build.log 2>&1
for args in "32 8 56 4 tc" "8 8 56 4 tc" "64 8 28 8 tc" "128 8 14 16 tc"; do timeout 120 python tools/prof_attn.py $args; done
GROUP_TIMEOUT=300 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py -k "tcgen05 or conv_igemm or stem_banded" 2>&1 | grep -c PASSED
grep -E "FAILED|ERROR" gpurun_out/bringup/custom.log | head
echo "=== bench batch 64"
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 --dump-ops gpurun_out/s7/ops_b64.jsonl > gpurun_out/s7/bench_b64.json 2> gpurun_out/s7/bench_b64.err; tail -c 2500 gpurun_out/s7/bench_b64.json; tail -3 gpurun_out/s7/bench_b64.err
