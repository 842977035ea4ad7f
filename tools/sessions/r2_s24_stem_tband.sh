#!/bin/bash
# round 2: temporal-band stem kernel -- correctness, A/B time against the banded stem at the headline shape, bench
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s24
mkdir -p $O
export ESF_NVCC_EXTRA="${EXTRA:-}"
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
[ "${TESTS:-1}" = 1 ] && timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "stem" > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -15 $O/pytest.log
for cfg in "1 0 8 64" "1 64 8 64" "1 6 8 64" "1 70 8 64"; do
set -- $cfg; tb=$1
BATCH=$4 ESF_STEM_TBAND=$1 ESF_STEM_TBAND_DBG=$2 ESF_STEM_TBAND_STAGES=$3 timeout 300 python - <<PY
import torch
from efficient_slowfast_b200.engine import Plan
g = torch.Generator().manual_seed(0)
import os
B, T = int(os.environ['BATCH']), 32
x = torch.randn(B, 3, T, 224, 224, device="cuda")
w = (torch.randn(8, 3, 5, 7, 7, generator=g) * 0.1).double()
b = (torch.randn(8, generator=g) * 0.1).double()
y = torch.empty(B, T, 112, 112, 8, dtype=torch.float16, device="cuda")
plan = Plan(torch.device("cuda"), "fp16")
plan.stem(x, y, w, b, (1, 2, 2), (2, 3, 3))
launch = plan.ops[-1] if hasattr(plan, "ops") else None
plan.launch_all(); torch.cuda.synchronize()
ref = y.clone()
import ctypes
from efficient_slowfast_b200 import runtime as rt
h = plan.handles[-1]
f = lambda: rt.check(rt.lib().esf_op_launch(h, None))
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print("tband=$tb dbg=$2 stages=$3 B=$4 %s: %.3f ms" % (plan.meta[-1]["label"], e0.elapsed_time(e1) / 20), "checksum %.6f" % y.float().abs().mean().item())
PY
done | tee $O/stem_ab.txt
if [ "${BENCH:-0}" = 1 ]; then ESF_STEM_TBAND=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-extra-configs > $O/bench.json 2> $O/bench.err; echo "bench rc $?"; fi
[ "${BENCH:-0}" = 1 ] && python - <<PY
import json
d = json.load(open("$O/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "parity_check") if k in d}, d.get("e2e", {}).get("value"))
PY
