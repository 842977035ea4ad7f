#!/bin/bash
# round 2: why are the fast-pathway 1x1x1 expansions that store into a concat slice 2x slower than their dense siblings?
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s23
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:igemm_kernel -s 110 -c 110 --csv --log-file $O/igemm_metrics.csv python bench.py --profile-mode --steps 1 --warmup 0 > $O/ncu.log 2>&1; tail -1 $O/ncu.log
