#!/bin/bash
# Runs the -m gpu tests in groups, each group in its own process (a device-side trap poisons only the CUDA context of
# its own process) with a timeout; one verdict line per test + full logs under gpurun_out/bringup/.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/bringup
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || { echo "BUILD FAILED"; tail -20 "$OUT/build.log"; exit 1; }
run_group() {
  local tag="$1"; shift
  local log="$OUT/$tag.log"
  timeout "${GROUP_TIMEOUT:-300}" python -m pytest "$@" -q -s -m gpu -rA > "$log" 2>&1
  local rc=$?
  echo "=== group $tag rc=$rc"
  grep -E "^(PASSED|FAILED|ERROR) " "$log" | sed 's/ - .*//' | head -60
  grep -E "DIAG|rel err|mbarrier wait timeout|illegal|misaligned|CUDA error|EsfError" "$log" | head -60
  if [ $rc -ne 0 ]; then grep -E "^E  " "$log" | head -25; fi
}
if [ $# -gt 0 ]; then
  run_group custom "$@"
else
  run_group simt tests/test_gpu_kernels.py -k "direct or stem or pool or eca or head or errors"
  run_group igemm tests/test_gpu_kernels.py -k "conv_igemm"
  run_group attn tests/test_gpu_kernels.py -k "attention"
  run_group model_s64 tests/test_gpu_model.py -k "s64 or stage_outputs or default_init or reload or pathway"
  run_group model_s224 tests/test_gpu_model.py -k "s224"
fi
