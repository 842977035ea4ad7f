#!/bin/bash
# round 2: 8-GPU weak scaling of the headline bench, e2e from FP32 host clips (slow-from-fast upload, H2D ceiling in the line)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s37
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
ls /sys/devices/system/node/ | tr '\n' ' '; echo; nvidia-smi topo -m 2>/dev/null | head -14
N=${1:-8}
for extra in ""; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 8 --warmup 3 $extra > $O/bench_n${N}${extra}.json 2> $O/bench_n${N}${extra}.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n${N}${extra}.json").read().strip().splitlines()[-1])
    print("N=$N $extra value %.0f ms %.2f | e2e %s | both %s | u8 %.0f" % (d["value"], d["ms_per_step"], json.dumps(d["e2e"]), json.dumps(d.get("e2e_both_pathways_uploaded")), d["e2e_uint8_frames"]["value"]))
except Exception as e:
    print("parse failed", e); print(open("$O/bench_n${N}${extra}.err").read()[-1500:])
PY
done
