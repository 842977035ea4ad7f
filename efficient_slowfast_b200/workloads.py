"""Synthetic workloads: the deterministic weight / input recipe and the named cases shared by the golden-vector
generator (tests/golden/make_golden.py, runs the REFERENCE in the build container), the tests, `__graft_entry__.smoke()`
and `bench.py` (run anywhere, no reference needed).  tests/golden/recipe.py and tests/helpers.py re-export this module.

Random-init parity is degenerate for this model family (SURVEY.md finding 5: ZERO_INIT_FINAL_BN and gamma = 0 zero
out every bottleneck and every attention branch), so parity runs use perturbed weights:
  * every tensor is drawn from a generator seeded by crc32(key) ^ seed  -> reproducible per key, on any machine;
  * conv weights ~ N(0, 2/fan_out) (c2_msra_fill), BN weight ~ U(0.5,1.5), BN bias ~ U(-0.2,0.2), attention
    gamma = 0.5, q/k/v biases ~ U(-0.1,0.1), ECA conv1d ~ U(-0.5,0.5), FC ~ N(0, 0.03), FC bias ~ U(-0.1,0.1);
  * the LAST BN of every bottleneck (`...branch2.c_bn`, the one the reference zero-initialises) gets
    weight ~ U(0.15,0.45) in the default "trained" recipe: the residual branch then changes the trunk by ~30 % per
    block (every conv matters) without the random network being chaotic.  With weight ~ U(0.5,1.5) there too (the
    "stress" recipe) a perturbation grows ~3x per stage -- 16-bit rounding noise of 2^-9 reaches 30 % of the
    activations at res5 in ANY implementation (measured on CPU by rounding the oracle's activations) --
    so that recipe is kept only as a stress fixture with an argmax check and a loose bound;
  * the position-attention query/key convs are scaled by 0.25 in the "trained" recipe: with c2_msra_fill weights the
    un-normalised logits q.k reach |s| ~ 200 on calibrated activations, the softmax is one-hot and the attention
    output flips between keys under perturbations of 1e-3 (a property of that random draw, not of an
    implementation); at |s| ~ 10 every key still matters and the branch is well conditioned;
  * Nonlocal blocks: theta / phi keep the c2_msra_fill scale (the d^-0.5 softmax logits then have std ~ 4: a few
    dozen keys matter; scaled down, the attention becomes uniform, the block output nearly constant and its BN divides
    by a variance of 0.02 -- measured with the reference); their final BN gets the U(0.15,0.45) weight of a residual
    branch;
  * BN running statistics are calibrated by the generator with train-mode forwards of the reference model and are
    STORED in the fixture (they cannot be regenerated without the reference).
"""
import os
import zlib

import numpy as np
import torch


def _gen(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def seeded_state_dict(template, seed=0, bn_stats=None, stress=False):
    """template: {key: tensor} from model.state_dict() (shapes only are used).  Returns a new FP32 state_dict."""
    out = {}
    for key, ref in template.items():
        shape = tuple(ref.shape)
        g = _gen(key, seed)
        is_bn = key.rsplit(".", 1)[0] + ".running_mean" in template
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.long)
        elif leaf in ("running_mean", "running_var"):
            if bn_stats is not None:
                t = torch.as_tensor(bn_stats[key]).reshape(shape).float().clone()
            else:
                t = torch.zeros(shape) if leaf == "running_mean" else torch.ones(shape)
        elif is_bn and leaf == "weight":
            t = torch.rand(shape, generator=g) + 0.5
            if (key.endswith("c_bn.weight") or key.endswith(".bn3.weight")
                    or ("_nonlocal" in key and key.endswith(".bn.weight"))) and not stress:
                t = (t - 0.5) * 0.3 + 0.15   # last BN of a residual branch (R50 bottleneck, ShuffleNet unit)
        elif is_bn and leaf == "bias":
            t = torch.rand(shape, generator=g) * 0.4 - 0.2
        elif leaf == "gamma":
            t = torch.full(shape, 0.5)
        elif leaf == "weight" and len(shape) == 5:
            fan_out = shape[0] * shape[2] * shape[3] * shape[4]
            t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
            if (".query_conv." in key or ".key_conv." in key) and not stress:
                t = t * 0.25   # keeps the un-scaled q.k logits at |s| ~ 10 instead of ~ 200 (see module docstring)
        elif leaf == "weight" and len(shape) == 3:
            t = torch.rand(shape, generator=g) - 0.5
        elif leaf == "weight" and len(shape) == 2:
            t = torch.randn(shape, generator=g) * 0.03
        elif leaf == "bias":
            t = torch.rand(shape, generator=g) * 0.2 - 0.1
        else:
            raise KeyError("recipe does not know how to fill %s %s" % (key, shape))
        out[key] = t
    return out


def seeded_clip(batch, frames, size, seed=1, channels=3):
    """x ~ N(0,1) of shape (B, C, T, S, S): the full-rate (fast pathway) clip; slow = pack_pathway_output(x)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, channels, frames, size, size, generator=g)


def pack_pathway_output(frames, alpha):
    """datasets/utils.py:73-112 (arch 'slowfast'): slow = frames[:, :, linspace(0, T-1, T//alpha).long()]."""
    if not alpha:                     # single-pathway archs (c2d / i3d / slow): datasets/utils.py:86-88
        return [frames]
    T = frames.shape[2]
    idx = torch.linspace(0, T - 1, T // alpha).long().to(frames.device)
    return [frames.index_select(2, idx).contiguous(), frames]


def sample_indices(numel, n=64, seed=7):
    g = torch.Generator().manual_seed(seed + numel % 1000003)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


# name -> (model, yaml under SlowFast/, cfg overrides, list of (tag, batch, frames, crop))
CASES = {
    "dual_r50": dict(
        model="SlowFastDualAttention", yaml="configs/Kinetics/SLOWFAST_DUAL_8x8_R50_stepwise_multigrid.yaml",
        opts=[], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64), ("s224", 1, 32, 224)]),
    "slowfast_r50": dict(
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml",
        opts=["MULTIGRID.SHORT_CYCLE", True], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64), ("s224", 1, 32, 224)]),
    "slowfast_r50_g2": dict(     # ResNeXt-style grouped 1x3x3 (RESNET.NUM_GROUPS = 2; resnet_helper.py:196-205)
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml",
        opts=["MULTIGRID.SHORT_CYCLE", True, "RESNET.NUM_GROUPS", 2], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64)]),
    "slowfast_r101": dict(       # RESNET.DEPTH 101: 23 blocks in res4 (video_model_builder.py:15)
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml",
        opts=["MULTIGRID.SHORT_CYCLE", True, "RESNET.DEPTH", 101], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64)]),
    "slowfast_r50_sigmoid": dict(    # MODEL.HEAD_ACT sigmoid (the multi-label configs, head_helper.py:189-196)
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml",
        opts=["MULTIGRID.SHORT_CYCLE", True, "MODEL.HEAD_ACT", "sigmoid", "MODEL.NUM_CLASSES", 157],
        calib=(2, 32, 96), inputs=[("s64", 2, 32, 64)]),
    "slowfast_r50_stress": dict(
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml", stress=True,
        opts=["MULTIGRID.SHORT_CYCLE", True], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64)]),
    "shufflenetv2_w05": dict(   # BASELINE configs[0]: SlowFastShuffleNetV2 width 0.5
        model="SlowFastShuffleNetV2", yaml="configs/Kinetics/SLOWFAST_SHUFFLENETV2_8x8_R50_stepwise_multigrid.yaml",
        opts=["SLOWFAST.WIDTH_MULTI", 0.5], calib=(2, 16, 112), inputs=[("s112", 2, 16, 112), ("s224", 1, 32, 224)]),
    "shufflenet_w2g3": dict(    # BASELINE configs[4]: SlowFastShuffleNet width 2.0 groups 3, Jester shape
        model="SlowFastShuffleNet", yaml="configs/Jester/SLOWFAST_SHUFFLENET_8x8_R50_stepwise_multigrid.yaml",
        opts=[], calib=(2, 16, 112), inputs=[("s112", 2, 16, 112), ("s64", 2, 16, 64)]),
    "mobilenetv2_w1": dict(     # BASELINE configs[3]: SlowFastMoibleNetV2 width 1.0
        model="SlowFastMoibleNetV2", yaml="configs/Kinetics/SLOWFAST_MOBILENETV2_8x8_R50_stepwise_multigrid.yaml",
        opts=["SLOWFAST.WIDTH_MULTI", 1.0], calib=(2, 16, 112), inputs=[("s112", 2, 16, 112), ("s224", 1, 32, 224)]),
    "ghostnet_w1": dict(        # BASELINE configs[3]: SlowFastGhostNet width 1.0 (224^2 x 32 frames has N = 100 352)
        model="SlowFastGhostNet", yaml="configs/Kinetics/SLOWFAST_GHOSTNET_8x8_R50_stepwise_multigrid.yaml",
        opts=["SLOWFAST.WIDTH_MULTI", 1.0], calib=(2, 16, 112),
        # s224 = the BASELINE shape; its golden needs make_golden.patch_large_n_attention (row-chunked reference)
        inputs=[("s112", 2, 16, 112), ("s64", 2, 16, 64), ("s224", 1, 32, 224)]),
    # section 8(f3): single-pathway ResNet.  The reference cannot build it with MULTIGRID.SHORT_CYCLE (its head gets two
    # pool sizes for one pathway, video_model_builder.py:584-586), so the crop is fixed per case.
    "i3d_r50": dict(            # I3D temporal kernels, (2,1,1) max-pool after res2
        model="ResNet", yaml="configs/Kinetics/I3D_8x8_R50.yaml", single=True,
        opts=[], calib=(2, 8, 224), inputs=[("s224", 1, 8, 224)]),
    "slow_r50": dict(
        model="ResNet", yaml="configs/Kinetics/SLOW_8x8_R50.yaml", single=True,
        opts=["DATA.CROP_SIZE", 64], calib=(2, 8, 64), inputs=[("s64", 2, 8, 64), ("s96", 1, 8, 96)]),
    # fully-convolutional inference: the head's AvgPool3d kernel ([4,2,2] / [32,2,2] at CROP_SIZE 64) is smaller than
    # the 3x3 feature map of a 96^2 clip -> Linear + softmax at 2x2 positions, then their mean (head_helper.py:218-220);
    # ("s96" of slow_r50 above is the single-pathway instance of the same path)
    "slowfast_r50_fcn": dict(
        model="SlowFast", yaml="configs/Kinetics/SLOWFAST_4x16_R50.yaml",
        opts=["DATA.CROP_SIZE", 64], calib=(2, 32, 96), inputs=[("s96", 1, 32, 96), ("s64", 2, 32, 64)]),
    # Nonlocal blocks after res3 blocks 1,3 and res4 blocks 1,3,5 (pool (1,2,2) on phi / g)
    "slow_nln_r50": dict(       # "dot_product" instantiation
        model="ResNet", yaml="configs/Kinetics/SLOW_NLN_8x8_R50.yaml", single=True,
        opts=["DATA.CROP_SIZE", 64], calib=(2, 8, 64), inputs=[("s64", 2, 8, 64)]),
    "i3d_nln_r50": dict(        # "softmax" instantiation, (2,1,1) max-pool after res2
        model="ResNet", yaml="configs/Kinetics/I3D_NLN_8x8_R50.yaml", single=True,
        opts=["DATA.CROP_SIZE", 96], calib=(2, 8, 96), inputs=[("s96", 2, 8, 96)]),
    # the fork's own product configs (configs/TIRED): grey-scale clips (one input channel), RESNET.DEPTH 18 with bottleneck
    # blocks (2, 2, 2, 2), ALPHA 8 with 16 frames (slow pathway: 2 frames), spatial strides (1, 1, 2, 2) so that a 112^2
    # crop reaches the 7 x 7 head pool of the default CROP_SIZE 224; s128 = its TEST_CROP_SIZE -> fully-convolutional head
    "dual_r18_gray": dict(
        model="SlowFastDualAttention", yaml="configs/TIRED/DUAL_TIRED_SLOWFAST_8x8_R18_HALF_112_GRAY.yaml", opts=[],
        channels=1, calib=(2, 16, 112), inputs=[("s112", 2, 16, 112), ("s128", 1, 16, 128)]),
    # MODEL.ARCH "fast" (video_model_builder.py:79-85, the fork's single-pathway arch with the fast pathway's temporal
    # kernels), DEPTH 18, WIDTH_PER_GROUP 16, grey-scale
    "fast_r18_gray": dict(
        model="ResNet", yaml="configs/TIRED/TIRED_FAST_NLN_8x8_R50_112.yaml", single=True, opts=[],
        channels=1, calib=(2, 16, 112), inputs=[("s112", 2, 16, 112)]),
    "dual_r50_stress": dict(
        model="SlowFastDualAttention", yaml="configs/Kinetics/SLOWFAST_DUAL_8x8_R50_stepwise_multigrid.yaml",
        stress=True, opts=[], calib=(2, 32, 96), inputs=[("s64", 2, 32, 64)]),
}


# ------------------------------------------------------------------------------------------------ named cases
GOLDEN_DIR = os.environ.get("ESF_GOLDEN_DIR") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")   # data fixtures (reference-made)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_cfg(name):
    """Our own cfg for a golden case (mirrors the YAML + overrides listed in CASES)."""
    from . import config as esf

    if name in ("dual_r50", "dual_r50_stress"):
        cfg = esf.slowfast_dual_8x8_r50_cfg()
    elif name in ("slowfast_r50", "slowfast_r50_stress"):
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.MULTIGRID.SHORT_CYCLE = True
    elif name == "slowfast_r101":
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.MULTIGRID.SHORT_CYCLE = True
        cfg.RESNET.DEPTH = 101
    elif name == "slowfast_r50_sigmoid":
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.MULTIGRID.SHORT_CYCLE = True
        cfg.MODEL.HEAD_ACT = "sigmoid"
        cfg.MODEL.NUM_CLASSES = 157
    elif name == "slowfast_r50_g2":
        cfg = esf.slowfast_4x16_r50_cfg()
        cfg.MULTIGRID.SHORT_CYCLE = True
        cfg.RESNET.NUM_GROUPS = 2
    elif name == "dual_r18_gray":       # configs/TIRED/DUAL_TIRED_SLOWFAST_8x8_R18_HALF_112_GRAY.yaml, model keys
        cfg = esf.slowfast_dual_8x8_r50_cfg()
        cfg.DATA._merge(dict(NUM_FRAMES=16, SAMPLING_RATE=2, TRAIN_CROP_SIZE=112, TEST_CROP_SIZE=128, CROP_SIZE=224,
                             INPUT_CHANNEL_NUM=[1, 1], MEAN=[0.45], STD=[0.225]))
        cfg.SLOWFAST._merge(dict(ALPHA=8, BETA_INV=8, FUSION_CONV_CHANNEL_RATIO=2, FUSION_KERNEL_SZ=7))
        cfg.RESNET._merge(dict(ZERO_INIT_FINAL_BN=True, WIDTH_PER_GROUP=64, NUM_GROUPS=1, DEPTH=18,
                               TRANS_FUNC="bottleneck_transform", STRIDE_1X1=False,
                               NUM_BLOCK_TEMP_KERNEL=[[2, 2]] * 4, SPATIAL_STRIDES=[[1, 1], [1, 1], [2, 2], [2, 2]],
                               SPATIAL_DILATIONS=[[1, 1]] * 4))
        cfg.NONLOCAL._merge(dict(LOCATION=[[[], []]] * 4, GROUP=[[1, 1]] * 4, INSTANTIATION="dot_product"))
        cfg.MODEL._merge(dict(NUM_CLASSES=3, ARCH="slowfast", MODEL_NAME="SlowFastDualAttention", DROPOUT_RATE=0.5))
        cfg.MULTIGRID.SHORT_CYCLE = False
    elif name == "fast_r18_gray":       # configs/TIRED/TIRED_FAST_NLN_8x8_R50_112.yaml, model keys
        cfg = esf.resnet_cfg("slow")
        cfg.DATA._merge(dict(NUM_FRAMES=16, SAMPLING_RATE=2, TRAIN_CROP_SIZE=112, TEST_CROP_SIZE=128, CROP_SIZE=224,
                             INPUT_CHANNEL_NUM=[1], MEAN=[0.45], STD=[0.225]))
        cfg.RESNET._merge(dict(ZERO_INIT_FINAL_BN=True, WIDTH_PER_GROUP=16, NUM_GROUPS=1, DEPTH=18,
                               TRANS_FUNC="bottleneck_transform", STRIDE_1X1=False, NUM_BLOCK_TEMP_KERNEL=[[2]] * 4,
                               SPATIAL_STRIDES=[[1], [1], [2], [2]], SPATIAL_DILATIONS=[[1]] * 4))
        cfg.NONLOCAL._merge(dict(LOCATION=[[[]]] * 4, GROUP=[[1]] * 4, INSTANTIATION="dot_product"))
        cfg.MODEL._merge(dict(NUM_CLASSES=3, ARCH="fast", MODEL_NAME="ResNet", DROPOUT_RATE=0.5))
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
    from .build import build_model

    cfg = case_cfg(name)
    cfg.ESF.PRECISION = precision
    torch.manual_seed(0)
    model = build_model(cfg)
    gold = load_golden(name)
    bn = {k[3:]: v for k, v in gold.items() if k.startswith("bn/")}
    sd = seeded_state_dict(model.state_dict(), seed=0, bn_stats=bn,
                                  stress=CASES[name].get("stress", False))
    model.load_state_dict(sd, strict=True)
    model.eval()
    return cfg, model, gold


def case_inputs(name, tag):
    for t, b, frames, crop in CASES[name]["inputs"]:
        if t == tag:
            cfg = case_cfg(name)
            alpha = 0 if CASES[name].get("single") else cfg.SLOWFAST.ALPHA
            return pack_pathway_output(seeded_clip(b, frames, crop, seed=1, channels=CASES[name].get("channels", 3)),
                                       alpha)
    raise KeyError(tag)


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
