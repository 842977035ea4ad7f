#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s54
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s -k "fcn or slow_r50" 2>&1 | grep -E "rel err|passed|failed|Error|error" | head -16
