#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s21
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_fp32_path.py -x -q -m gpu -k "attention or golden or stage" > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra-configs --no-cpu-baseline > $O/bench_b64.json 2> $O/bench_b64.err
python - <<PY
import json
d=json.loads(open("$O/bench_b64.json").read().strip().splitlines()[-1])
print("bench", round(d["value"],1), round(d["ms_per_step"],3), d["parity_check"]["rel_err"], d["parity_check"]["ok"])
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
