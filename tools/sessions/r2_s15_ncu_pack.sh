#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s15
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
cat > $O/pk.py <<PY
import sys; sys.path.insert(0, ".")
import torch
from efficient_slowfast_b200 import runtime as rt
L = rt.lib()
B, T = 64, 32
x = torch.randn(B, 3, T, 224, 224, device="cuda")
pitch, lpad, _ = rt.stem_geometry(224, 3, 7, 2, 3)
xp = torch.empty(B, T, 224, pitch, dtype=torch.float16, device="cuda")
for _ in range(3):
    rt.check(L.esf_stem_pack(x.data_ptr(), B, 3, T, 224, 224, pitch, lpad, rt.F16, xp.data_ptr(), None))
torch.cuda.synchronize()
PY
for m in 0 1; do ESF_STEM_PACK_SMEM=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_pack -s 2 -c 1 -o $O/prof_pack_smem$m python $O/pk.py > $O/ncu$m.log 2>&1; tail -1 $O/ncu$m.log; done
