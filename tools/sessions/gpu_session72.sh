#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s72
mkdir -p $O
for L in 0 1; do
ESF_NVCC_EXTRA=-DESF_ATTN_P_IN_S=$L python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > $O/build_$L.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attn or attention" > $O/tests_$L.log 2>&1; tail -1 $O/tests_$L.log | sed "s/^/p_in_s=$L /"; grep -E "^FAILED|Error|timeout" $O/tests_$L.log | head -5
for D in 8 32; do
  timeout 300 python tools/prof_attn.py $D 8 56 4 tc 2 > $O/prof_${L}_$D.log 2>&1; tail -1 $O/prof_${L}_$D.log | cut -c1-150 | sed "s/^/p_in_s=$L d=$D /"
done
done
