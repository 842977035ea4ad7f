#!/bin/bash
# round 2: compute-sanitizer over the temporal-band stem kernel (kernel tests + one model golden)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s29
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
{
echo "# compute-sanitizer over stem_tband_kernel (tools/sessions/r2_s29_sanitizer_stem.sh, one B200)"
echo "## memcheck, kernel tests"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "temporal_band and not 224" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
echo "## racecheck, kernel tests (fast stem geometry)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "temporal_band and 2-3-8-64 and fp16" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -8
echo "## memcheck, model golden through the new stem"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "test_model_matches_reference_golden and slowfast_r50 and s64" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
} | tee $O/sanitizer.txt
