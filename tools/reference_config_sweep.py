"""Container-only: build every YAML under the reference's SlowFast/configs with the reference's build_model and with this
package's, and compare the state_dict schema (keys, shapes) and the seeded initial weights.
    python tools/reference_config_sweep.py > profiles/r2_reference_config_sweep.txt"""
import sys, os, glob, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tests/golden')
import torch
from oracle import ref_shim
import efficient_slowfast_b200 as esf
root='/root/reference/SlowFast/configs'
ok=bad_ref=bad_ours=mismatch=0
for y in sorted(glob.glob(root+'/**/*.yaml', recursive=True)):
    rel=os.path.relpath(y, '/root/reference/SlowFast')
    try:
        cfg=ref_shim.get_cfg(rel, [])
        if cfg.DETECTION.ENABLE: print("SKIP detection", rel); continue
        torch.manual_seed(0)
        ref=ref_shim.build_reference_model(cfg)
    except Exception as e:
        bad_ref+=1; print("REF-FAIL", rel, type(e).__name__, str(e)[:80]); continue
    try:
        ours_cfg=esf.get_cfg(); ours_cfg.merge_from_file(y); ours_cfg.NUM_GPUS=0
        torch.manual_seed(0)
        ours=esf.build_model(ours_cfg)
    except Exception as e:
        bad_ours+=1; print("OURS-FAIL", rel, type(e).__name__, str(e)[:120]); continue
    a={k:tuple(v.shape) for k,v in ref.state_dict().items()}; b={k:tuple(v.shape) for k,v in ours.state_dict().items()}
    if a!=b:
        mismatch+=1; print("MISMATCH", rel, len(a), len(b), list(set(a)^set(b))[:4])
    else:
        same=all(torch.equal(ref.state_dict()[k], ours.state_dict()[k]) for k in a)
        ok+=1; print("OK", rel, len(a), "bit-identical init" if same else "init differs")
print("ok",ok,"ref-fail",bad_ref,"ours-fail",bad_ours,"mismatch",mismatch)
