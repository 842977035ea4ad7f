#!/bin/bash
# round 2, first pass: MUFU.EX2 f32 vs f16 / f16x2 rate, full GPU suite, smoke launch list (are our kernels inside the
# driver's first 1000 launches?), headline bench with parity_check / configs / H2D ceiling
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s1
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ex2 tools/ubench/ex2.cu && /tmp/ex2 > $O/ubench_ex2.txt 2>&1
cat $O/ubench_ex2.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -15 $O/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2_s1/smoke_launches.csv")) if len(r)>5]
H=rows[0]; ik=H.index("Kernel Name")
c=collections.Counter(r[ik].split("(")[0][:60] for r in rows[1:])
print(len(rows)-1,"launches"); [print(n,k) for k,n in c.most_common(25)]
PY
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --dump-ops $O/ops_b64.jsonl > $O/bench_b64.json 2> $O/bench_b64.err
tail -1 $O/bench_b64.json | cut -c1-600; tail -3 $O/bench_b64.err
