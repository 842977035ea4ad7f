#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s27
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --dump-ops $O/ops_mobilenetv2.jsonl --model SlowFastMoibleNetV2 --batch 8 --frames 32 --crop 224 > $O/b1.json 2> $O/b1.err; tail -2 $O/b1.err
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --dump-ops $O/ops_shufflenet.jsonl --model SlowFastShuffleNet --batch 8 --frames 16 --crop 112 > $O/b2.json 2> $O/b2.err; tail -2 $O/b2.err
