#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s79
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or direct" 2>&1 | tail -15
echo "== tma march (search)"; timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma.log
echo "== hs=28"; ESF_DW_HS=28 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma_hs28.log
for c in 2 3 6 8 9; do echo "== ch8=$c"; ESF_DW_CH8=$c timeout 300 python tools/prof_dwconv.py 3 0,1,4,8 2>&1 | tee $O/prof_tma_ch$c.log; done
