"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the efficient two-stream models (same rules as slowfast_oracle.py).

Functional, state_dict-driven restatement (torch CPU) of the eval-mode forward of

  * SlowFastShuffleNetV2  (SlowFast/slowfast/models/custom_video_model_builder.py:448-617)
  * SlowFastShuffleNet    (custom_video_model_builder.py:620-789)

Pinned by tests/test_oracle_vs_reference.py (bit-exact stage outputs against the real reference, build container only)
and by tests/golden/*.npz produced by the reference itself.
"""
import torch
import torch.nn.functional as F

from .slowfast_oracle import _bn, _conv, fuse_fast_and_slow

V2_TABLE = {0.25: [-1, 24, 32, 64, 128, 1024], 0.5: [-1, 24, 48, 96, 192, 1024], 1.0: [-1, 24, 116, 240, 464, 1024],
            1.5: [-1, 24, 176, 352, 704, 1024], 2.0: [-1, 24, 224, 496, 976, 2048]}   # custom_video_model_builder.py:470-481
V1_TABLE = {1: [24, 144, 288, 567], 2: [24, 200, 400, 800], 3: [24, 240, 480, 960], 4: [24, 272, 544, 1088],
            8: [24, 384, 768, 1536]}                                                   # custom_video_model_builder.py:643-652
STAGE_REPEATS = [4, 8, 4]


def channel_shuffle(x, groups):
    """shufflenetv2_helper.py:32-43 / shufflenet_helper.py:24-34."""
    b, c, t, h, w = x.shape
    return x.view(b, groups, c // groups, t, h, w).permute(0, 2, 1, 3, 4, 5).contiguous().view(b, c, t, h, w)


def stem_3x3x3_pool(x, sd, p):
    """shufflenetv2_stem / shufflenet_stem (stem_helper.py:237-246, 274-284): Conv3d 3x3x3 s(1,2,2) p1 -> BN -> ReLU ->
    MaxPool3d k3 s(1,2,2) p1 (pads T with -inf)."""
    x = F.relu(_bn(_conv(x, sd, p + ".0", stride=(1, 2, 2), padding=(1, 1, 1)), sd, p + ".1"))
    return F.max_pool3d(x, kernel_size=3, stride=(1, 2, 2), padding=1)


# ------------------------------------------------------------------------------------------------ ShuffleNetV2
def v2_unit(x, sd, p, stride):
    """InvertedResidual.forward (shufflenetv2_helper.py:104-112)."""
    def branch2(z):
        z = F.relu(_bn(_conv(z, sd, p + ".banch2.0"), sd, p + ".banch2.1"))
        c = z.shape[1]
        z = _bn(_conv(z, sd, p + ".banch2.3", stride=(1, stride, stride), padding=1, groups=c), sd, p + ".banch2.4")
        return F.relu(_bn(_conv(z, sd, p + ".banch2.5"), sd, p + ".banch2.6"))

    if stride == 1:
        c = x.shape[1] // 2
        out = torch.cat((x[:, :c], branch2(x[:, c:])), 1)
    else:
        cin = x.shape[1]
        z = _bn(_conv(x, sd, p + ".banch1.0", stride=(1, stride, stride), padding=1, groups=cin), sd, p + ".banch1.1")
        z = F.relu(_bn(_conv(z, sd, p + ".banch1.2"), sd, p + ".banch1.3"))
        out = torch.cat((z, branch2(x)), 1)
    return channel_shuffle(out, 2)


def basic_head(feats, sd, p, act="softmax", return_logits=False):
    """Tail shared by the efficient heads (head_helper.py:470-486, 540-557, 594-609): global avg pool per pathway -> cat
    -> NTHWC -> Linear (classifier.1) -> softmax(dim=4) -> mean."""
    pooled = [F.avg_pool3d(x, x.shape[-3:]) for x in feats]
    x = torch.cat(pooled, 1).permute(0, 2, 3, 4, 1)
    logits = F.linear(x, sd[p + ".classifier.1.weight"].to(x.dtype), sd[p + ".classifier.1.bias"].to(x.dtype))
    y = torch.softmax(logits, dim=4) if act == "softmax" else torch.sigmoid(logits)
    y = y.mean([1, 2, 3]).reshape(x.shape[0], -1)
    return (y, logits) if return_logits else y


def slowfast_shufflenetv2_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFastShuffleNetV2.forward (custom_video_model_builder.py:604-617), eval mode."""
    alpha, beta = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV
    so = V2_TABLE[cfg.SLOWFAST.WIDTH_MULTI]
    fo = [c // beta for c in so]
    sd = {k: v.detach().to("cpu") for k, v in sd.items()}
    xs = [t.detach().to("cpu", dtype) for t in inputs]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val] if isinstance(val, list) else val.clone()

    xs = [stem_3x3x3_pool(xs[p], sd, "s1.pathway%d_stem" % p) for p in range(2)]
    tap("s1", xs)
    xs = fuse_fast_and_slow(xs, sd, "s1_fuse", alpha)
    tap("s1_fuse", xs)
    for i in range(1, 4):
        stage = "s%d" % (i + 1)
        nxt = []
        for p, chans in enumerate((so, fo)):
            x = xs[p]
            pre = "%s.pathway%d_channel_%d.features" % (stage, p, chans[i + 1])
            for u in range(STAGE_REPEATS[i - 1]):
                x = v2_unit(x, sd, "%s.%d" % (pre, u), 2 if u == 0 else 1)
            nxt.append(x)
        xs = nxt
        tap(stage, xs)
        xs = fuse_fast_and_slow(xs, sd, stage + "_fuse", alpha)
        tap(stage + "_fuse", xs)
    feats = []
    for p in range(2):
        q = "head.pathway%d_conv1x1x1.0" % p
        feats.append(F.relu(_bn(_conv(xs[p], sd, q + ".0"), sd, q + ".1")))
    y, logits = basic_head(feats, sd, "head", cfg.MODEL.HEAD_ACT, return_logits=True)
    tap("logits", logits)
    tap("head", y)
    return y


# ------------------------------------------------------------------------------------------------ ShuffleNet (v1)
def v1_unit(x, sd, p, stride, groups):
    """Bottleneck.forward (shufflenet_helper.py:75-84); grouping of conv1 per :48 (g = 1 iff in_planes == 24)."""
    g1 = 1 if x.shape[1] == 24 else groups
    out = F.relu(_bn(_conv(x, sd, p + ".conv1", groups=g1), sd, p + ".bn1"))
    out = channel_shuffle(out, groups)
    c = out.shape[1]
    out = _bn(_conv(out, sd, p + ".conv2", stride=(1, stride, stride), padding=1, groups=c), sd, p + ".bn2")
    out = _bn(_conv(out, sd, p + ".conv3", groups=groups), sd, p + ".bn3")
    if stride == 2:
        sc = F.avg_pool3d(_conv(x, sd, p + ".shortcut.0"), kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
        return F.relu(torch.cat([out, sc], 1))
    return F.relu(out + x)


def slowfast_shufflenet_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFastShuffleNet.forward (custom_video_model_builder.py:776-789), eval mode."""
    alpha, beta, groups = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV, cfg.SLOWFAST.GROUPS
    so = [int(i * cfg.SLOWFAST.WIDTH_MULTI) for i in V1_TABLE[groups]]
    fo = [c // beta for c in so]
    sd = {k: v.detach().to("cpu") for k, v in sd.items()}
    xs = [t.detach().to("cpu", dtype) for t in inputs]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val] if isinstance(val, list) else val.clone()

    xs = [stem_3x3x3_pool(xs[p], sd, "s1.pathway%d_stem" % p) for p in range(2)]
    tap("s1", xs)
    xs = fuse_fast_and_slow(xs, sd, "s1_fuse", alpha)
    tap("s1_fuse", xs)
    for i in range(3):
        stage = "s%d" % (i + 2)
        nxt = []
        for p, chans in enumerate((so, fo)):
            x = xs[p]
            pre = "%s.pathway%d_channel_%d.features" % (stage, p, chans[i + 1])
            for u in range(STAGE_REPEATS[i]):
                x = v1_unit(x, sd, "%s.%d" % (pre, u), 2 if u == 0 else 1, groups)
            nxt.append(x)
        xs = nxt
        tap(stage, xs)
        xs = fuse_fast_and_slow(xs, sd, stage + "_fuse", alpha)
        tap(stage + "_fuse", xs)
    y, logits = basic_head(xs, sd, "head", cfg.MODEL.HEAD_ACT, return_logits=True)
    tap("logits", logits)
    tap("head", y)
    return y


FORWARDS = {
    "SlowFastShuffleNetV2": slowfast_shufflenetv2_forward,
    "SlowFastShuffleNet": slowfast_shufflenet_forward,
}
