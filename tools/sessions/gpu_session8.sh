#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s8
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s8/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -o gpurun_out/s8/prof_attn_tc_d32 python tools/prof_attn.py 32 8 56 3 tc 1 > gpurun_out/s8/ncu1.log 2>&1; tail -2 gpurun_out/s8/ncu1.log
