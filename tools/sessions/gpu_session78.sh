#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s78
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise or direct" 2>&1 | tail -15
echo "== tma march"; timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma.log
echo "== tma march wbt=1"; ESF_DW_WBT=1 timeout 300 python tools/prof_dwconv.py 3 2>&1 | tee $O/prof_tma_wbt1.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dwconv -s 1 -c 1 -o $O/dw_tma_c32 -f python tools/prof_dwconv.py 1 2 > $O/ncu_c32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dwconv -s 1 -c 1 -o $O/dw_tma_c144 -f python tools/prof_dwconv.py 1 0 > $O/ncu_c144.log 2>&1
