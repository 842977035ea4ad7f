#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s49
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s -k "row_softmax or transpose16 or nonlocal" 2>&1 | grep -v "^$" | tail -15
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "nln" 2>&1 | grep -E "rel err|passed|failed|Error|error" | head -12
