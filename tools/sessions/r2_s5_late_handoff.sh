#!/bin/bash
# round 2: late P hand-over (tcgen05.wait::st + arrive half way through the NEXT tile's exponentials) together with the
# second P buffer -- round 1 measured the late hand-over alone at -12 % because the single P buffer then stalled
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s5
mkdir -p $O
ESF_NVCC_EXTRA="-DESF_ATTN_LATE_HANDOFF=1" python -c "from efficient_slowfast_b200 import _build; print(_build.build(force=True))" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
export ESF_NVCC_EXTRA="-DESF_ATTN_LATE_HANDOFF=1"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" > $O/pytest_attn.log 2>&1; echo "pytest (late, 1 P buffer) rc $?"; tail -2 $O/pytest_attn.log
ESF_ATTN_PDBL=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_tcgen05" > $O/pytest_attn2.log 2>&1; echo "pytest (late, 2 P buffers) rc $?"; tail -2 $O/pytest_attn2.log
for pd in 0 1; do for d in 8 16 32; do
  echo -n "LATE=1 PDBL=$pd  "; ESF_ATTN_PDBL=$pd timeout 120 python tools/prof_attn.py $d 8 56 16 tc 5 2>&1 | tail -1
done; done | tee $O/late_ab.txt
