#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s32
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
P="python tools/prof_conv.py"
ESF_PROF_GROUPS=96 $P 32 32 112 112 96 96 3 3 3 1 2 2 0
ESF_PROF_GROUPS=144 $P 32 32 56 56 144 144 3 3 3 1 1 1 0
ESF_PROF_GROUPS=384 $P 64 32 14 14 384 384 3 3 3 1 1 1 0
ESF_PROF_GROUPS=144 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 1 -c 1 -o $O/prof_dw144 $P 32 32 56 56 144 144 3 3 3 1 1 1 0 2 > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
