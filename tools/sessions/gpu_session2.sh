#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/s2
bash tools/gpu_bringup.sh tests/test_gpu_model.py 2>&1 | tail -60
echo "=== bench batch 8"
timeout 600 python bench.py --batch 8 --steps 3 --warmup 3 > gpurun_out/s2/bench_b8.json 2> gpurun_out/s2/bench_b8.err; tail -c 3000 gpurun_out/s2/bench_b8.json; tail -3 gpurun_out/s2/bench_b8.err
echo "=== bench batch 64"
timeout 900 python bench.py --batch 64 --steps 5 --warmup 3 > gpurun_out/s2/bench_b64.json 2> gpurun_out/s2/bench_b64.err; tail -c 3000 gpurun_out/s2/bench_b64.json; tail -3 gpurun_out/s2/bench_b64.err
echo "=== ncu launch list (batch 8, eager)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"igemm_kernel|attn_|stem_conv|pool3d|eca_|head_|conv_direct" -s 143 -c 143 --csv --log-file gpurun_out/s2/launches_b8.csv python bench.py --batch 8 --steps 1 --profile-mode > gpurun_out/s2/ncu_b8.log 2>&1; tail -2 gpurun_out/s2/ncu_b8.log; wc -l gpurun_out/s2/launches_b8.csv
