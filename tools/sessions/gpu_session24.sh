#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s24
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
# launch list of the bench command (profile mode = same plan launched eagerly so every kernel is visible to ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -2 $O/launches.log
# full capture of the dominant kernel (attention d=32 = 2nd attn_tc_v2 launch of a step), taken in the 2nd step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 3 -c 1 -o $O/prof_bench_attn_d32 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn.log 2>&1; tail -2 $O/ncu_attn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 2 -c 1 -o $O/prof_bench_attn_d8 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn8.log 2>&1; tail -2 $O/ncu_attn8.log
