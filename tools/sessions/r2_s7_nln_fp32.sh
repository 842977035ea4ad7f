#!/bin/bash
# round 2: Non-local blocks in the FP32-accurate plan + FP32-path regression
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s7
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fp32_path.py -q -s -m gpu > $O/pytest_fp32.log 2>&1; echo "pytest fp32 rc $?"
grep -E "fp32: rel err|passed|failed|Error|error|assert" $O/pytest_fp32.log | head -40
