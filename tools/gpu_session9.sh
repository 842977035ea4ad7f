#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/s9
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s9/build.log 2>&1
GROUP_TIMEOUT=300 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py -k "plumbing or pool or head" 2>&1 | tail -12
GROUP_TIMEOUT=600 bash tools/gpu_bringup.sh tests/test_gpu_model.py -k "shufflenet" 2>&1 | tail -60
