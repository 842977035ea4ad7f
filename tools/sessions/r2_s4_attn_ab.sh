#!/bin/bash
# round 2: attention v2 experiments -- second P buffer (d < 32) and independent Q.K^T issue of the two query tiles
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s4
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" > $O/pytest_attn.log 2>&1; echo "pytest attention rc $?"; tail -3 $O/pytest_attn.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  for d in 8 32; do
    echo -n "PDBL=$1 QKASYNC=$2  "; ESF_ATTN_PDBL=$1 ESF_ATTN_QKASYNC=$2 timeout 120 python tools/prof_attn.py $d 8 56 16 tc 5 2>&1 | tail -1
  done
done | tee $O/attn_ab.txt
echo -n "d=64  QKASYNC n/a (v1 kernel): "; timeout 120 python tools/prof_attn.py 64 8 28 16 tc 5 2>&1 | tail -1 | tee -a $O/attn_ab.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err
python - <<PY
import json
d=json.loads(open("$O/bench_b64.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], d["parity_check"], d["kernel_breakdown"]["attention"])
PY
