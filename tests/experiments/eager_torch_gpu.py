"""GPU experiment (test infrastructure, not collected by pytest): the eager PyTorch forward of the headline model ON THE
B200 -- the torch ops the reference itself launches (cuDNN conv3d, cuBLAS bmm, eager softmax / BN / pools), driven
by the oracle's functional restatement with its tensors on the GPU.  This is the "recompiled library kernels" bar of
SURVEY.md section 8(d), next to the CUDA path of this repo on the same clips.  The reference materialises the N x N
affinity of every position attention (2.5 GB FP32 per clip and stage at 224^2); the oracle's row-chunked form does the
same arithmetic without it, so this baseline is, if anything, kinder to eager PyTorch.

    python tests/experiments/eager_torch_gpu.py [batch] [crop] [fp32|tf32|bf16]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for q in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, q)

import helpers  # noqa: E402
import recipe  # noqa: E402
from oracle import slowfast_oracle as O  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    crop = int(sys.argv[2]) if len(sys.argv) > 2 else 224
    mode = sys.argv[3] if len(sys.argv) > 3 else "tf32"
    case = sys.argv[4] if len(sys.argv) > 4 else "dual_r50"
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    cfg, model, _ = helpers.case_model_and_weights(case)
    cfg.DATA.CROP_SIZE = crop
    alpha = 0 if recipe.CASES[case].get("single") else cfg.SLOWFAST.ALPHA
    frames = cfg.DATA.NUM_FRAMES
    O.DEVICE = "cuda"
    sd = {k: v.cuda() for k, v in model.state_dict().items()}
    xs = [t.cuda() for t in recipe.pack_pathway_output(recipe.seeded_clip(batch, frames, crop, seed=1), alpha)]
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32

    def step():
        return O.forward(cfg, sd, xs, dtype=dtype)

    y = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        y = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the CUDA path of this repo on the same clips
    m = model.cuda().eval()
    with torch.no_grad():
        z = m(xs)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            z = m(xs)
        e1.record()
        torch.cuda.synchronize()
    ms_ours = e0.elapsed_time(e1) / steps
    print(json.dumps({"experiment": "eager torch ops on the GPU (oracle restatement, row-chunked attention)",
                      "case": case, "batch": batch, "frames": frames, "crop": crop, "math": mode,
                      "eager_ms_per_step": ms, "eager_clips_per_s": batch / ms * 1e3,
                      "esf_ms_per_step": ms_ours, "esf_clips_per_s": batch / ms_ours * 1e3,
                      "speedup": ms / ms_ours,
                      "rel_err_esf_vs_eager": helpers.rel_err(z.float().cpu(), y.float().cpu()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "wall_s": time.perf_counter() - t0}))


if __name__ == "__main__":
    main()
