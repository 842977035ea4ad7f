#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s88
mkdir -p $O
ESF_NVCC_EXTRA=-DESF_DW_ROTATE_MOV python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or direct" 2>&1 | tail -5
timeout 300 python tools/prof_dwconv.py 5 2>&1 | tee $O/prof.log
