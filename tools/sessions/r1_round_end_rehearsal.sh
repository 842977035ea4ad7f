#!/bin/bash
# round-end rehearsal: full GPU suite, smoke(), headline bench (both arms) as the driver runs them
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s100
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err
tail -1 $O/bench_b64.json | cut -c1-300
