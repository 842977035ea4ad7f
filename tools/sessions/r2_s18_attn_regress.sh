#!/bin/bash
# round 2: is the attention kernel of this round (pdbl / qk_async / split code paths added, off by default) slower than
# round 1's?  Same box, same process flow: time both builds with tools/prof_attn.py at batch 64.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/r2_s18
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || { tail -5 $O/build.log; exit 1; }
for rep in 1 2; do for d in 8 32; do echo -n "round-2 kernel: "; timeout 120 python tools/prof_attn.py $d 8 56 64 tc 5 2>&1 | tail -1; done; done | tee $O/new.txt
cp efficient_slowfast_b200/csrc/esf_attn_tc.cu $O/esf_attn_tc_r2.cu.bak
cp tools/ubench/esf_attn_tc_r1.cu.txt efficient_slowfast_b200/csrc/esf_attn_tc.cu
# the round-1 file lacks the entry points added this round: append stubs so that the library links and loads
cat >> efficient_slowfast_b200/csrc/esf_attn_tc.cu <<'CU'
extern "C" int64_t esf_attn_tc_vlo_bytes(int32_t, int32_t, int32_t) { return -1; }
extern "C" int esf_attn_tc_pack_vlo(const float*, int32_t, int32_t, int32_t, void*, void*) { return -1; }
extern "C" int esf_attn_tc_create_split(const void*, const void*, int32_t, int32_t, int32_t, int32_t, int32_t, float, const float*, const float*, int32_t, const esf_view*, esf_op**) { return -1; }
CU
python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > $O/build_old.log 2>&1 || { tail -5 $O/build_old.log; }
for rep in 1 2; do for d in 8 32; do echo -n "round-1 kernel: "; timeout 120 python tools/prof_attn.py $d 8 56 64 tc 5 2>&1 | tail -1; done; done | tee $O/old.txt
# back to the round-2 source; attention tests with the knobs off and on (the knob paths are a separate instantiation)
cp $O/esf_attn_tc_r2.cu.bak efficient_slowfast_b200/csrc/esf_attn_tc.cu
python -c "from efficient_slowfast_b200 import _build; _build.build(force=True)" > /dev/null 2>&1
for k in "0 0" "1 1"; do set -- $k
  ESF_ATTN_PDBL=$1 ESF_ATTN_QKASYNC=$2 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fp32_path.py -x -q -m gpu -k "attention" 2>&1 | tail -1
done
