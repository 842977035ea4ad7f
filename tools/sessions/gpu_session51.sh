#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s51
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s -k "nln" 2>&1 | grep -E "rel err|passed|failed|Error|error" | head -12
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
