#!/bin/bash
# round 2: ncu --set full of the HBM-bound helper kernels of the headline step (ECA, pool, stem pack, head pool, attn pack)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s12
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:eca_|pool3d|stem_pack|head_pool|attn_tc_pack' -s 17 -c 17 \
  -o $O/prof_small python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_small.log 2>&1; tail -2 $O/ncu_small.log
