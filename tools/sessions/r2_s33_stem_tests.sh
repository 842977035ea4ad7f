#!/bin/bash
# round 2: stem kernel tests (fixed + randomised geometries)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s33
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stem" > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -15 $O/pytest.log
