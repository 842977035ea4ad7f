#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s6
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s6/build.log 2>&1
for args in "32 8 56 4 tc" "32 8 56 4 mma" "8 8 56 4 tc" "8 8 56 4 mma" "64 8 28 8 tc" "64 8 28 8 mma"; do timeout 120 python tools/prof_attn.py $args; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -s 1 -c 1 -o gpurun_out/s6/prof_attn_tc_d32 python tools/prof_attn.py 32 8 56 1 tc 1 > gpurun_out/s6/ncu1.log 2>&1; tail -2 gpurun_out/s6/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 1 -c 1 -o gpurun_out/s6/prof_attn_mma_d32 python tools/prof_attn.py 32 8 56 1 mma 1 > gpurun_out/s6/ncu2.log 2>&1; tail -2 gpurun_out/s6/ncu2.log
ls -la gpurun_out/s6
