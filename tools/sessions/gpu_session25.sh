#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s25
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -2 $O/launches.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 600 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
