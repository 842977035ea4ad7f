#!/bin/bash
# ncu launch list of one MobileNetV2 step (batch 128) with the depthwise march / padded pointwise kernels
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s95
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale|dwconv|pw_small|row_softmax|transpose16|group_mean|frames_|global_mean'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file $O/launches_mnv2.csv python bench.py --model SlowFastMoibleNetV2 --batch 128 --frames 32 --crop 224 --profile-mode --steps 1 --warmup 0 > $O/launches_mnv2.log 2>&1; tail -1 $O/launches_mnv2.log
timeout 600 python bench.py --model SlowFastGhostNet --batch 32 --frames 32 --crop 224 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_ghostnet.json 2> $O/bench_ghostnet.err; tail -c 300 $O/bench_ghostnet.json
