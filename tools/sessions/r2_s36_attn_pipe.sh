#!/bin/bash
# round 2: software-pipelined softmax loop of the position attention (ESF_ATTN_PIPE=1) -- kernel tests and time
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s36
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
{
for pipe in ${PIPES:-0 1}; do
  echo "== ESF_ATTN_PIPE=$pipe"
  for d in 8 32; do ESF_ATTN_PIPE=$pipe timeout 120 python tools/prof_attn.py $d 8 56 16 tc 5 2>&1 | tail -${TAIL:-1}; done
done
ESF_ATTN_PIPE=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -${TTAIL:-3}
} | tee $O/pipe.txt
