#!/bin/bash
# A/B of P aliased onto its S tile in the v2 attention kernel
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s71
mkdir -p $O
for L in 0 1; do
ESF_NVCC_EXTRA=-DESF_ATTN_P_IN_S=$L python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > $O/build_$L.log 2>&1
for D in 8 32; do
  timeout 300 python tools/prof_attn.py $D 8 56 16 tc 5 2>&1 | tail -1 | sed "s/^/p_in_s=$L /"
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn or attention" 2>&1 | tail -1 | sed "s/^/p_in_s=$L /"
done
python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_b64.json 2> $O/bench_b64.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s71/bench_b64.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], {k:v['ms'] for k,v in list(d['kernel_breakdown'].items())[:3]})
PY
