#!/bin/bash
# round 2: is the d = 32 attention bound by the Q.K^T issuer?  Timing with fewer Q.K^T MMAs per tile (wrong logits)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s30
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
for n in 0 4 2 1; do
  echo "d=32 qk steps $n (0 = all 6):"; ESF_ATTN_QK_STEPS_DBG=$n python tools/prof_attn.py 32 8 56 16 tc 5 2>&1 | tail -1
done | tee $O/qk_steps.txt
for n in 0 1; do
  echo "d=8 qk steps $n (0 = all 2):"; ESF_ATTN_QK_STEPS_DBG=$n python tools/prof_attn.py 8 8 56 16 tc 5 2>&1 | tail -1
done | tee -a $O/qk_steps.txt
