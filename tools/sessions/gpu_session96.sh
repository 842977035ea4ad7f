#!/bin/bash
# memcheck / racecheck of the kernels added this session (depthwise march, padded pointwise, grouped pool, slice route)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s96
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or pool3d or pointwise_tiny or unaligned" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" $O/memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise_vector_kernel and fp16" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck.log | tail -3
