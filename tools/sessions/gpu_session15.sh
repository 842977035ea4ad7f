#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s15
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
P="python tools/prof_conv.py"
$P 64 8 56 56 64 256 1 1 1 1 1 1 1
$P 64 8 56 56 64 256 1 1 1 1 1 1 0
$P 64 8 56 56 256 64 1 1 1 1 1 1 0
$P 64 32 56 56 8 8 1 3 3 1 1 1 0
$P 64 32 56 56 8 8 1 3 3 1 1 1 0 5 fp16 0
$P 64 32 56 56 8 32 1 1 1 1 1 1 1
$P 64 32 56 56 8 32 1 1 1 1 1 1 1 5 fp16 0
$P 64 32 56 56 32 8 3 1 1 1 1 1 0
$P 64 32 56 56 32 8 3 1 1 1 1 1 0 5 fp16 0
$P 64 32 28 28 16 16 1 3 3 1 1 1 0
$P 64 32 28 28 16 16 1 3 3 1 1 1 0 5 fp16 0
$P 64 32 28 28 16 64 1 1 1 1 1 1 1
$P 64 32 28 28 16 64 1 1 1 1 1 1 1 5 fp16 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 1 -c 1 -o $O/prof_c64to256_res $P 64 8 56 56 64 256 1 1 1 1 1 1 1 2 > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 1 -c 1 -o $O/prof_wfold_8to32_res $P 64 32 56 56 8 32 1 1 1 1 1 1 1 2 > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log
