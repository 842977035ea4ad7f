#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s28
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; tail -12 $O/gpu_tests.log
run() { name=$1; shift; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dump-ops $O/ops_$name.jsonl "$@" > $O/bench_$name.json 2> $O/bench_$name.err; python - "$O/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(d['config']['workload'], '| %.1f clips/s  %.2f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
    print('   ', {k:(v['ms'],v['launches']) for k,v in d['kernel_breakdown'].items()})
except Exception as e:
    print('FAILED', sys.argv[1], e)
PY
tail -2 $O/bench_$name.err; }
run shufflenetv2 --model SlowFastShuffleNetV2 --batch 64 --frames 32 --crop 224
run mobilenetv2 --model SlowFastMoibleNetV2 --batch 128 --frames 32 --crop 224
run shufflenet --model SlowFastShuffleNet --batch 256 --frames 16 --crop 112
run ghostnet --model SlowFastGhostNet --batch 32 --frames 32 --crop 224
