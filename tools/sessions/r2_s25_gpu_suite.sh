#!/bin/bash
# round 2: whole GPU suite + smoke + short bench after a kernel change
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s25
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_b64.json 2> $O/bench_b64.err
tail -1 $O/bench_b64.json | cut -c1-300; tail -3 $O/bench_b64.err
