#!/bin/bash
# round 2: the fork's TIRED configs (grey-scale, R18, ARCH fast), grouped conv -- 16-bit and FP32 plans
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s9
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_model.py -q -s -m gpu -k "gray or g2" > $O/pytest_new.log 2>&1; echo "pytest rc $?"
grep -E "rel err|passed|failed|Error|error|assert" $O/pytest_new.log | head -40
timeout 900 python -m pytest tests/test_gpu_fp32_path.py -q -s -m gpu -k "gray" > $O/pytest_new32.log 2>&1; echo "pytest fp32 rc $?"
grep -E "rel err|passed|failed|Error|error|assert" $O/pytest_new32.log | head -20
