#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s44
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale|dwconv|pw_small'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -1 $O/launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 3 -c 1 -o $O/prof_bench_attn_d32 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn.log 2>&1; tail -1 $O/ncu_attn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 2 -c 1 -o $O/prof_bench_attn_d8 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn8.log 2>&1; tail -1 $O/ncu_attn8.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; tail -c 300 $O/bench_b64.json
