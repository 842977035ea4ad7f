#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s20
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
ESF_ATTN_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 1 -c 1 -o $O/prof_attn_v3_d32 python tools/prof_attn.py 32 8 56 3 tc 1 > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
