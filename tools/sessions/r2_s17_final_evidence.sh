#!/bin/bash
# round 2 evidence run (one B200): round-end sequence as the driver runs it + FP32 plan + ncu launch list
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s32
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err
tail -1 $O/bench_b64.json | cut -c1-400; tail -3 $O/bench_b64.err
for b in 16 32; do
  timeout 600 python bench.py --precision fp32 --batch $b --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_fp32_b$b.json 2> $O/bench_fp32_b$b.err
  tail -1 $O/bench_fp32_b$b.json | cut -c1-200
done
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale|dwconv|pw_small|row_softmax|transpose16|group_mean|frames_|p32_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -1 $O/launches.log
