#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s81
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python bench.py --model SlowFastMoibleNetV2 --batch 128 --frames 32 --crop 224 --steps 3 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_mobilenetv2.jsonl > $O/bench_mnv2.json 2> $O/bench_mnv2.err
timeout 900 python bench.py --model SlowFastGhostNet --batch 32 --frames 32 --crop 224 --steps 3 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_ghostnet.jsonl > $O/bench_ghost.json 2> $O/bench_ghost.err
ls -la $O
