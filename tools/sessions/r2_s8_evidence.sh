#!/bin/bash
# round 2 evidence: compute-sanitizer over the kernels added this round, ncu launch list of one bench step, ncu --set full
# of the dominant kernel (d = 32 attention at batch 64), reference arm
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s8
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
SEL="post_split or conv_fp32 or pool_eca or attention_fp32"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fp32_path.py -q -m gpu -k "$SEL" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fp32_path.py -q -m gpu -k "pool_eca or attention_fp32 or post_split" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/sanitizer_racecheck.log | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_model.py -q -m gpu -k "forward_fast or short_last_batch" > $O/sanitizer_memcheck_model.log 2>&1; echo "memcheck(model) rc $?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitizer_memcheck_model.log | tail -3
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale|dwconv|pw_small|row_softmax|transpose16|group_mean|frames_|p32_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -1 $O/launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 3 -c 1 -o $O/prof_bench_attn_d32 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn.log 2>&1; tail -1 $O/ncu_attn.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 500 $O/bench_reference.json; echo
