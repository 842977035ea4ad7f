#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s46
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "i3d or slow_r50 or perform_test" 2>&1 | grep -E "rel err of probs|passed|failed|Error" | head -12
