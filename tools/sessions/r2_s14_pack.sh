#!/bin/bash
# round 2: stem pack kernel variants -- correctness + time of the two packs of the headline step
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s14
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_frames.py -x -q -m gpu -k "stem or frames or pool" > $O/pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/pytest.log
for smem in 0 1; do
ESF_STEM_PACK_SMEM=$smem python - <<PY
import torch, ctypes
from efficient_slowfast_b200 import runtime as rt
L = rt.lib()
for (B, T) in ((64, 32), (64, 8)):
    x = torch.randn(B, 3, T, 224, 224, device="cuda")
    pitch, lpad, _ = rt.stem_geometry(224, 3, 7, 2, 3)
    xp = torch.empty(B, T, 224, pitch, dtype=torch.float16, device="cuda")
    f = lambda: rt.check(L.esf_stem_pack(x.data_ptr(), B, 3, T, 224, 224, pitch, lpad, rt.F16, xp.data_ptr(), None))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("smem=$smem T=%d: %.3f ms  %.0f GB/s" % (T, ms, (x.numel() * 4 + xp.numel() * 2) / ms / 1e6))
PY
done | tee $O/pack_ab.txt
