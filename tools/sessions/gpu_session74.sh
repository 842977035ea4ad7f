#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s74
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or direct" 2>&1 | tail -5
echo "== old tile kernel"; ESF_DW_MARCH=0 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_old.log
echo "== march VEC8"; timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_march_v8.log
echo "== march VEC4"; ESF_DW_VEC=4 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_march_v4.log
echo "== march VEC8 hs=28"; ESF_DW_HS=28 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_march_v8_hs28.log
