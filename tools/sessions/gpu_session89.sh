#!/bin/bash
# 8-GPU line of the headline bench (launched as the driver does)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s89
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -1 $O/bench_n$N.json | cut -c1-400
