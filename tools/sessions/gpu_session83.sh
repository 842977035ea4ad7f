#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s83
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dwconv -s 1 -c 1 -o $O/dw_tma_c32 -f python tools/prof_dwconv.py 1 2 > $O/ncu_c32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dwconv -s 1 -c 1 -o $O/dw_tma_c144 -f python tools/prof_dwconv.py 1 0 > $O/ncu_c144.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dwconv -s 1 -c 1 -o $O/dw_tma_c96_s2 -f python tools/prof_dwconv.py 1 1 > $O/ncu_c96.log 2>&1
timeout 300 python tools/prof_dwconv.py 5 2>&1 | tee $O/prof_final.log
ESF_DW_MARCH=0 timeout 300 python tools/prof_dwconv.py 5 2>&1 | tee $O/prof_old.log
