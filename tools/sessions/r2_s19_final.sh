#!/bin/bash
# round 2, final evidence run (one B200): round-end sequence as the driver runs it, FP32 plan, ncu launch list, and the
# attention A/B against round 1's source
set -u
cd "$(dirname "$0")/../.."
bash tools/sessions/r2_s17_final_evidence.sh
sed -i 's|O=gpurun_out/r2_s18|O=gpurun_out/r2_s19_ab|' tools/sessions/r2_s18_attn_regress.sh
bash tools/sessions/r2_s18_attn_regress.sh 2>&1 | grep -E "kernel:|passed|failed"
