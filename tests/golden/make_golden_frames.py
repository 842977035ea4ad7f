"""Generate tests/golden/frames_input.npz by running the UNMODIFIED reference loader functions (tensor_normalize,
pack_pathway_output of SlowFast/slowfast/datasets/utils.py, imported from /root/reference through oracle/ref_shim.py)
on seeded uint8 frames.  Build-container only; the fixture is committed.

    python tests/golden/make_golden_frames.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

CASES = {   # name -> (alpha, arch, reverse, mean, std)
    "a4": (4, "slowfast", False, [0.45, 0.45, 0.45], [0.225, 0.225, 0.225]),
    "a8_rev_imagenet": (8, "slowfast", True, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
    "single": (4, "slow", False, [0.45, 0.45, 0.45], [0.225, 0.225, 0.225]),
}


def seeded_frames(seed=3, shape=(2, 16, 12, 20, 3)):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)


def main():
    U = ref_shim.reference_dataset_utils()
    frames = seeded_frames()
    out = {"frames": frames.numpy()}
    for name, (alpha, arch, reverse, mean, std) in CASES.items():
        cfg = ref_shim.get_cfg()
        cfg.SLOWFAST.ALPHA = alpha
        cfg.MODEL.ARCH = arch
        cfg.DATA.REVERSE_INPUT_CHANNEL = reverse
        clips = []
        for b in range(frames.shape[0]):   # datasets/kinetics.py:231-248
            f = U.tensor_normalize(frames[b], mean, std)
            f = f.permute(3, 0, 1, 2)
            clips.append(U.pack_pathway_output(cfg, f))
        for i in range(len(clips[0])):
            out["%s/%d" % (name, i)] = torch.stack([c[i] for c in clips]).numpy()
    np.savez_compressed(os.path.join(HERE, "frames_input.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
