#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s99
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "depthwise" 2>&1 | tail -3
timeout 300 python tools/prof_dwconv.py 5 2>&1 | tail -13
for M in SlowFastMoibleNetV2:128:32:224 SlowFastShuffleNet:256:16:112; do
IFS=: read model batch frames crop <<< "$M"
timeout 900 python bench.py --model $model --batch $batch --frames $frames --crop $crop --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$model.json 2> $O/bench_$model.err; python - $O/bench_$model.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), {k:v['ms'] for k,v in list(d['kernel_breakdown'].items())[:5]})
PY
done
