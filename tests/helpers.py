"""Shared test helpers (no reference needed)."""
import os

import numpy as np
import torch

import recipe  # tests/golden/recipe.py

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_cfg(name):
    """Our own cfg for a golden case (mirrors the YAML + overrides listed in recipe.CASES)."""
    import efficient_slowfast_b200 as esf

    if name in ("dual_r50", "dual_r50_stress"):
        cfg = esf.slowfast_dual_8x8_r50_cfg()
    elif name in ("slowfast_r50", "slowfast_r50_stress"):
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.MULTIGRID.SHORT_CYCLE = True
    elif name == "shufflenetv2_w05":
        cfg = esf.slowfast_shufflenetv2_cfg(0.5)
    elif name == "mobilenetv2_w1":
        cfg = esf.slowfast_mobilenetv2_cfg(1.0)
    elif name == "ghostnet_w1":
        cfg = esf.slowfast_ghostnet_cfg(1.0)
    elif name == "shufflenet_w2g3":
        cfg = esf.slowfast_shufflenet_cfg(2.0, 3)
    elif name == "i3d_r50":
        cfg = esf.resnet_cfg("i3d")
    elif name == "slow_r50":
        cfg = esf.resnet_cfg("slow")
        cfg.DATA.CROP_SIZE = 64
    elif name == "slowfast_r50_fcn":
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.DATA.CROP_SIZE = 64
    elif name == "slow_nln_r50":
        cfg = esf.resnet_cfg("slow", nln=True)
        cfg.DATA.CROP_SIZE = 64
    elif name == "i3d_nln_r50":
        cfg = esf.resnet_cfg("i3d", nln=True)
        cfg.DATA.CROP_SIZE = 96
    else:
        raise KeyError(name)
    cfg.NUM_GPUS = 0
    return cfg


def case_model_and_weights(name, precision="fp16"):
    """(cfg, model on CPU with the seeded + calibrated golden weights loaded)."""
    import efficient_slowfast_b200 as esf

    cfg = case_cfg(name)
    cfg.ESF.PRECISION = precision
    torch.manual_seed(0)
    model = esf.build_model(cfg)
    gold = load_golden(name)
    bn = {k[3:]: v for k, v in gold.items() if k.startswith("bn/")}
    sd = recipe.seeded_state_dict(model.state_dict(), seed=0, bn_stats=bn,
                                  stress=recipe.CASES[name].get("stress", False))
    model.load_state_dict(sd, strict=True)
    model.eval()
    return cfg, model, gold


def case_inputs(name, tag):
    for t, b, frames, crop in recipe.CASES[name]["inputs"]:
        if t == tag:
            cfg = case_cfg(name)
            alpha = 0 if recipe.CASES[name].get("single") else cfg.SLOWFAST.ALPHA
            return recipe.pack_pathway_output(recipe.seeded_clip(b, frames, crop, seed=1), alpha)
    raise KeyError(tag)


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
