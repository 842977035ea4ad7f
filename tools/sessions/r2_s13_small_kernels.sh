#!/bin/bash
# round 2: reworked HBM-bound helpers (stem pack through shared memory, packed 16-bit max-pool, head pool grid): tests + step
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s16
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err
python - <<PY
import json
d=json.loads(open("$O/bench_b64.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), round(d["ms_per_step"],3), d["parity_check"]["rel_err"], d["parity_check"]["ok"], "e2e", round(d["e2e"]["value"],1))
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
