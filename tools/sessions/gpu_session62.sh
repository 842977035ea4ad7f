#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s62
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 1 -c 1 -o $O/prof_igemm_64to32 python tools/prof_conv.py 64 8 56 56 64 32 1 1 1 1 1 1 0 1 fp16 0 > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel -s 1 -c 1 -o $O/prof_igemm_64to256_res python tools/prof_conv.py 64 8 56 56 64 256 1 1 1 1 1 1 1 1 fp16 0 > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log
