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


# ------------------------------------------------------------------------------------------------ MobileNetV2
MBV2_SETTING = [[1, 16, 1, (1, 1, 1)], [6, 24, 2, (1, 2, 2)], [6, 32, 3, (1, 2, 2)], [6, 64, 4, (1, 2, 2)],
                [6, 96, 3, (1, 1, 1)], [6, 160, 3, (1, 2, 2)], [6, 320, 1, (1, 1, 1)]]   # custom_video_model_builder.py:1029-1047


def mbv2_unit(x, sd, p, stride, expand_ratio, oup):
    """InvertedResidual.forward (mobilenetv2_helper.py:64-68)."""
    inp = x.shape[1]
    if expand_ratio == 1:
        y = F.relu6(_bn(_conv(x, sd, p + ".conv.0", stride=stride, padding=1, groups=inp), sd, p + ".conv.1"))
        y = _bn(_conv(y, sd, p + ".conv.3"), sd, p + ".conv.4")
    else:
        y = F.relu6(_bn(_conv(x, sd, p + ".conv.0"), sd, p + ".conv.1"))
        y = F.relu6(_bn(_conv(y, sd, p + ".conv.3", stride=stride, padding=1, groups=y.shape[1]), sd, p + ".conv.4"))
        y = _bn(_conv(y, sd, p + ".conv.6"), sd, p + ".conv.7")
    return x + y if (tuple(stride) == (1, 1, 1) and inp == oup) else y


def slowfast_mobilenetv2_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFastMoibleNetV2.forward (custom_video_model_builder.py:1264-1285), eval mode."""
    alpha, beta, wm = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV, cfg.SLOWFAST.WIDTH_MULTI
    sd = {k: v.detach().to("cpu") for k, v in sd.items()}
    xs = [t.detach().to("cpu", dtype) for t in inputs]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val] if isinstance(val, list) else val.clone()

    def stage(xs, name, rows):
        out = []
        for p in range(2):
            x = xs[p]
            pre = "%s.pathway%d_channel_%d.features" % (name, p, rows[0][1])
            u = 0
            for t, c, n, s in rows:
                oup = int(c * wm) if p == 0 else int(c * wm // beta)
                for i in range(n):
                    x = mbv2_unit(x, sd, "%s.%d" % (pre, u), s if i == 0 else (1, 1, 1), t, oup)
                    u += 1
            out.append(x)
        tap(name, out)
        return out

    def fuse(xs, name):
        xs = fuse_fast_and_slow(xs, sd, name, alpha)
        tap(name, xs)
        return xs

    L = MBV2_SETTING
    xs = [F.relu6(_bn(_conv(xs[p], sd, "s1.pathway%d_stem.features.0" % p, stride=(1, 2, 2), padding=(1, 1, 1)), sd,
                      "s1.pathway%d_stem.features.1" % p)) for p in range(2)]
    tap("s1", xs)
    xs = fuse(stage(xs, "s2", L[0:2]), "s3_fuse")
    xs = fuse(stage(xs, "s4", L[2:3]), "s4_fuse")
    xs = fuse(stage(xs, "s5", L[3:4]), "s5_fuse")
    xs = stage(xs, "s6", L[4:5])
    xs = fuse(stage(xs, "s7", L[5:6]), "s7_fuse")
    xs = stage(xs, "s8", L[6:])
    feats = []
    for p in range(2):
        q = "head.pathway%d_conv1x1x1" % p
        feats.append(F.relu6(_bn(_conv(xs[p], sd, q + ".0"), sd, q + ".1")))
    y, logits = basic_head(feats, sd, "head", cfg.MODEL.HEAD_ACT, return_logits=True)
    tap("logits", logits)
    tap("head", y)
    return y


# ------------------------------------------------------------------------------------------------ GhostNet
GHOST_STAGES = [   # k, t, c, SE, s   (custom_video_model_builder.py:813-845)
    [[3, 16, 16, 0, 1]],
    [[3, 48, 24, 0, 2], [3, 72, 24, 0, 1]],
    [[5, 72, 40, 0.25, 2], [5, 120, 40, 0.25, 1]],
    [[3, 240, 80, 0, 2], [3, 200, 80, 0, 1], [3, 184, 80, 0, 1], [3, 184, 80, 0, 1], [3, 480, 112, 0.25, 1],
     [3, 672, 112, 0.25, 1]],
    [[5, 672, 160, 0.25, 2], [5, 960, 160, 0, 1], [5, 960, 160, 0.25, 1], [5, 960, 160, 0, 1], [5, 960, 160, 0.25, 1]],
]


def make_divisible(v, divisor, min_value=None):
    """ghostnet_helper.py:11-24."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def ghost_module(x, sd, p, oup, relu):
    """GhostModule.forward (ghostnet_helper.py:95-99): the primary conv is 1x1x1 in every use (kernel_size=1), the cheap
    operation a depthwise 3x3x3; concat then slice to `oup` channels."""
    x1 = _bn(_conv(x, sd, p + ".primary_conv.0"), sd, p + ".primary_conv.1")
    if relu:
        x1 = F.relu(x1)
    x2 = _bn(_conv(x1, sd, p + ".cheap_operation.0", padding=1, groups=x1.shape[1]), sd, p + ".cheap_operation.1")
    if relu:
        x2 = F.relu(x2)
    return torch.cat([x1, x2], 1)[:, :oup]


def ghost_bottleneck(x, sd, p, mid, out, k, stride, se_ratio):
    """GhostBottleneck.forward (ghostnet_helper.py:149-163); SqueezeExcite :46-52 with the hard-sigmoid gate :27-31."""
    residual = x
    y = ghost_module(x, sd, p + ".ghost1", mid, True)
    if stride > 1:
        y = _bn(_conv(y, sd, p + ".conv_dw", stride=(1, stride, stride), padding=(0, (k - 1) // 2, (k - 1) // 2),
                      groups=y.shape[1]), sd, p + ".bn_dw")
    if se_ratio:
        g = y.mean(dim=(2, 3, 4), keepdim=True)
        g = F.relu(_conv(g, sd, p + ".se.conv_reduce"))
        g = _conv(g, sd, p + ".se.conv_expand")
        y = y * (F.relu6(g + 3.0) / 6.0)
    y = ghost_module(y, sd, p + ".ghost2", out, False)
    if (p + ".shortcut.0.weight") in sd:
        r = _bn(_conv(residual, sd, p + ".shortcut.0", stride=(1, stride, stride),
                      padding=(0, (k - 1) // 2, (k - 1) // 2), groups=residual.shape[1]), sd, p + ".shortcut.1")
        residual = _bn(_conv(r, sd, p + ".shortcut.2"), sd, p + ".shortcut.3")
    return y + residual


def slowfast_ghostnet_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFastGhostNet.forward (custom_video_model_builder.py:1009-1026), eval mode.  The head's `act` is ReLU
    (head_helper.py:653), so the output is ReLU(logits)."""
    alpha, beta, wm = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV, cfg.SLOWFAST.WIDTH_MULTI
    sd = {k: v.detach().to("cpu") for k, v in sd.items()}
    xs = [t.detach().to("cpu", dtype) for t in inputs]
    slow = [[[c[0], make_divisible(c[1] * wm, 4), make_divisible(c[2] * wm, 4), c[3], c[4]] for c in st]
            for st in GHOST_STAGES]
    fast = [[[c[0], make_divisible(c[1] * wm // beta, 4), make_divisible(c[2] * wm // beta, 4), c[3], c[4]] for c in st]
            for st in GHOST_STAGES]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val] if isinstance(val, list) else val.clone()

    xs = [F.relu(_bn(_conv(xs[p], sd, "s0.pathway%d_stem.0" % p, stride=(1, 2, 2), padding=1), sd,
                     "s0.pathway%d_stem.1" % p)) for p in range(2)]
    tap("s0", xs)
    for i in range(5):
        name = "s%d" % (i + 1)
        out = []
        for p, rows in enumerate((slow[i], fast[i])):
            x = xs[p]
            pre = "%s.pathway%d_channel_%d.features" % (name, p, rows[-1][2])
            for u, (k, t, c, se, s) in enumerate(rows):
                x = ghost_bottleneck(x, sd, "%s.%d" % (pre, u), make_divisible(t, 2), make_divisible(c, 2), k, s, se)
            out.append(x)
        xs = out
        tap(name, xs)
        if i < 4:
            xs = fuse_fast_and_slow(xs, sd, name + "_fuse", alpha)
            tap(name + "_fuse", xs)
    pooled = []
    for p, tag in enumerate(("slow", "fast")):
        x = F.relu(_bn(_conv(xs[p], sd, "head.stage5_conv_%s.conv" % tag), sd, "head.stage5_conv_%s.bn1" % tag))
        x = F.avg_pool3d(x, x.shape[-3:])
        pooled.append(F.relu(_conv(x, sd, "head.conv_head_%s" % tag)))
    x = torch.cat(pooled, 1).permute(0, 2, 3, 4, 1)
    logits = F.linear(x, sd["head.classifier.1.weight"].to(x.dtype), sd["head.classifier.1.bias"].to(x.dtype))
    y = F.relu(logits).mean([1, 2, 3]).reshape(x.shape[0], -1)
    tap("logits", logits)
    tap("head", y)
    return y


FORWARDS = {
    "SlowFastGhostNet": slowfast_ghostnet_forward,
    "SlowFastShuffleNetV2": slowfast_shufflenetv2_forward,
    "SlowFastShuffleNet": slowfast_shufflenet_forward,
    "SlowFastMoibleNetV2": slowfast_mobilenetv2_forward,
    "SlowFastMobileNetV2": slowfast_mobilenetv2_forward,
}
