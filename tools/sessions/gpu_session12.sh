#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s12
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s12/build.log 2>&1
GROUP_TIMEOUT=300 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py -k "tcgen05" 2>&1 | grep -E "FAILED|PASSED|timeout|=== " | head -30
for args in "32 8 56 4 tc" "8 8 56 4 tc" "64 8 28 8 tc"; do timeout 120 python tools/prof_attn.py $args; done
