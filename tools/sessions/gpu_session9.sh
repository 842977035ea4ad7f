#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s9
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s9/build.log 2>&1
GROUP_TIMEOUT=300 bash tools/gpu_bringup.sh tests/test_gpu_kernels.py -k "plumbing or generic or head" 2>&1 | tail -8
GROUP_TIMEOUT=900 bash tools/gpu_bringup.sh tests/test_gpu_model.py -k "shufflenet or mobilenet or ghostnet" 2>&1 | tail -60
