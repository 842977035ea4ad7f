#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s10
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s10/build.log 2>&1
GROUP_TIMEOUT=600 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py 2>&1 | grep -E "^=== |FAILED|ERROR|DIAG" | head -30
grep -c PASSED gpurun_out/bringup/custom.log
GROUP_TIMEOUT=1200 bash tools/gpu_bringup.sh tests/test_gpu_model.py -k "golden" 2>&1 | grep -E "FAILED|rel err of probs|=== " | head -70
