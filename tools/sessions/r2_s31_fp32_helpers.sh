#!/bin/bash
# round 2: FP32-plan helper kernels after the 32-bit index / per-block gate rework: tests + bench at batch 16
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s31
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_fp32_path.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest.log
for b in 16 32; do
timeout 600 python bench.py --precision fp32 --batch $b --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_fp32_b$b.json 2> $O/bench_fp32_b$b.err
python - <<PY
import json
x = json.loads(open("$O/bench_fp32_b$b.json").read().strip().splitlines()[-1])
print("batch $b: %.1f clips/s %.2f ms  parity %.2e" % (x["value"], x["ms_per_step"], x["parity_check"]["rel_err"]))
print({k: v["ms"] for k, v in x["kernel_breakdown"].items()})
PY
done
