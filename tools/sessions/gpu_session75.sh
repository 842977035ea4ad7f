#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s76
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or direct" 2>&1 | tail -15
echo "== tma march"; timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma.log
echo "== tma march hs=28"; ESF_DW_HS=28 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma_hs28.log
