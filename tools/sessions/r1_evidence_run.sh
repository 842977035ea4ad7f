#!/bin/bash
# Round-1 evidence run: default bench (with CPU baseline), reference arm, ncu launch list + full captures of the two
# big attention launches, per-op times, the other BASELINE configs and the section-8(f3) models.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s56
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err; tail -c 400 $O/bench_b64.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json; echo
K='regex:igemm_kernel|attn_|stem_|pool3d|eca_|head_|conv_direct|shuffle_|eltwise_|channel_scale|dwconv|pw_small|row_softmax|transpose16|group_mean|frames_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file $O/launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/launches.log 2>&1; tail -1 $O/launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 3 -c 1 -o $O/prof_bench_attn_d32 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn.log 2>&1; tail -1 $O/ncu_attn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_v2 -s 2 -c 1 -o $O/prof_bench_attn_d8 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_attn8.log 2>&1; tail -1 $O/ncu_attn8.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc_pack -s 5 -c 1 -o $O/prof_bench_attn_pack_d32 python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu_pack.log 2>&1; tail -1 $O/ncu_pack.log
run() { name=$1; shift; timeout 900 python bench.py "$@" --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err; python - $O/bench_$name.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],1), 'u8', round(d.get('e2e_uint8_frames',{}).get('value',0),1))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
run slowfast_r50 --model SlowFast --batch 32
run shufflenetv2 --model SlowFastShuffleNetV2 --batch 64
run mobilenetv2 --model SlowFastMoibleNetV2 --batch 128
run ghostnet --model SlowFastGhostNet --batch 32
run shufflenet --model SlowFastShuffleNet --batch 256 --frames 16 --crop 112
run slow_nln_r50 --case slow_nln_r50 --frames 8 --batch 64
run i3d_nln_r50 --case i3d_nln_r50 --frames 8 --batch 64
run slow_r50 --case slow_r50 --frames 8 --batch 64
run i3d_r50 --case i3d_r50 --frames 8 --batch 64
