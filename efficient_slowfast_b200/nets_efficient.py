"""The efficient two-stream models of the reference behind the same registry / nn.Module surface:

  SlowFastShuffleNetV2   (SlowFast/slowfast/models/custom_video_model_builder.py:448-617)
  SlowFastShuffleNet     (custom_video_model_builder.py:620-789)

Same contract as nets_resnet.py: the module tree only holds parameters under the reference's state_dict keys
(`s#.pathway#_channel_<C>.features.#.banch{1,2}.#.*`, `...conv{1,2,3}.weight`, ...); the
arithmetic is a launch plan of libesf_b200 kernels.  These backbones have channel counts that are not multiples of 8
(27, 54, 6, ...), so most of their layers run on the CUDA-core kernels (`esf_conv_direct`, `esf_pool3d`,
`esf_shuffle_concat`); layers whose views happen to be 16-byte addressable go to the tensor-core implicit GEMM
automatically (engine.Plan.conv).
"""
import math

import torch
import torch.nn as nn

from . import runtime as rt
from .build import MODEL_REGISTRY
from .engine import fold_conv_bn
from .nets_resnet import FuseFastAndSlow, _Holder, _PlannedModel, get_norm, init_weights


def _stage_init(module):
    """The per-stage `_initialize_weights` of the reference helpers (e.g. shufflenetv2_helper.py:283-297): conv weights
    ~ N(0, 2 / (k_t k_h k_w C_out)), BN weight 1 / bias 0, Linear N(0, 0.01).  Run at the same points of construction
    as in the reference so that a given torch seed produces the same weights."""
    for m in module.modules():
        if isinstance(m, nn.Conv3d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm3d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
        elif isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 0.01)
            m.bias.data.zero_()


def _conv_out(n, k, s, p, d=1):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


class _EfficientBase(_PlannedModel):
    """Plan-emission helpers shared by the efficient models."""

    def _emit_conv(self, plan, x, conv, bn, act, y=None, res=None, name=None):
        """conv (+BN) (+act) of one nn.Conv3d; allocates the output when `y` is None."""
        B, T, H, W, _ = x.shape
        k, s, p, d = conv.kernel_size, conv.stride, conv.padding, conv.dilation
        To, Ho, Wo = (_conv_out(T, k[0], s[0], p[0], d[0]), _conv_out(H, k[1], s[1], p[1], d[1]),
                      _conv_out(W, k[2], s[2], p[2], d[2]))
        if y is None:
            y = plan.act(B, To, Ho, Wo, conv.out_channels, name=name)
        w, b = fold_conv_bn(conv.weight, conv.bias, bn)
        plan.conv(x, y, w, b, stride=tuple(s), padding=tuple(p), dilation=tuple(d), groups=conv.groups, act=act, res=res)
        return y

    def _emit_stem(self, plan, seq, x_nc, dst, act):
        """Sequential(Conv3d, BN, ReLU|ReLU6[, MaxPool3d]) on the FP32 clip; writes into `dst` (a concat slice)."""
        conv, bn = seq[0], seq[1]
        B, _, T, H, W = x_nc.shape
        k, s, p = conv.kernel_size, conv.stride, conv.padding
        To, Ho, Wo = _conv_out(T, k[0], s[0], p[0]), _conv_out(H, k[1], s[1], p[1]), _conv_out(W, k[2], s[2], p[2])
        w, b = fold_conv_bn(conv.weight, conv.bias, bn)
        pool = seq[3] if len(seq) > 3 else None
        if pool is None:
            y = plan.act(B, To, Ho, Wo, conv.out_channels)
            plan.stem(x_nc, y, w, b, tuple(s), tuple(p), act=act)
            plan.shuffle_concat(y, None, 1, dst)   # plain copy into the slice (groups = 1)
            return
        y = plan.act(B, To, Ho, Wo, conv.out_channels)
        plan.stem(x_nc, y, w, b, tuple(s), tuple(p), act=act)
        pk = pool.kernel_size if isinstance(pool.kernel_size, (tuple, list)) else (pool.kernel_size,) * 3
        ps = pool.stride if isinstance(pool.stride, (tuple, list)) else (pool.stride,) * 3
        pp = pool.padding if isinstance(pool.padding, (tuple, list)) else (pool.padding,) * 3
        plan.pool(y, dst, tuple(pk), tuple(ps), tuple(pp))

    @staticmethod
    def _pooled_shape(T, H, W, seq):
        conv = seq[0]
        k, s, p = conv.kernel_size, conv.stride, conv.padding
        T, H, W = _conv_out(T, k[0], s[0], p[0]), _conv_out(H, k[1], s[1], p[1]), _conv_out(W, k[2], s[2], p[2])
        if len(seq) > 3:
            pool = seq[3]
            pk = pool.kernel_size if isinstance(pool.kernel_size, (tuple, list)) else (pool.kernel_size,) * 3
            ps = pool.stride if isinstance(pool.stride, (tuple, list)) else (pool.stride,) * 3
            pp = pool.padding if isinstance(pool.padding, (tuple, list)) else (pool.padding,) * 3
            T, H, W = _conv_out(T, pk[0], ps[0], pp[0]), _conv_out(H, pk[1], ps[1], pp[1]), _conv_out(W, pk[2], ps[2], pp[2])
        return T, H, W


# ================================================================================================ ShuffleNetV2
class InvertedResidual(_Holder):
    """shufflenetv2_helper.py:46-112 (sic: 'banch')."""

    def __init__(self, inp, oup, stride):
        super().__init__()
        self.stride = stride
        assert stride in [1, 2]
        oup_inc = oup // 2
        if self.stride == 1:
            self.banch2 = nn.Sequential(
                nn.Conv3d(oup_inc, oup_inc, 1, 1, 0, bias=False), nn.BatchNorm3d(oup_inc), nn.ReLU(inplace=True),
                nn.Conv3d(oup_inc, oup_inc, 3, (1, stride, stride), 1, groups=oup_inc, bias=False),
                nn.BatchNorm3d(oup_inc),
                nn.Conv3d(oup_inc, oup_inc, 1, 1, 0, bias=False), nn.BatchNorm3d(oup_inc), nn.ReLU(inplace=True))
        else:
            self.banch1 = nn.Sequential(
                nn.Conv3d(inp, inp, 3, (1, stride, stride), 1, groups=inp, bias=False), nn.BatchNorm3d(inp),
                nn.Conv3d(inp, oup_inc, 1, 1, 0, bias=False), nn.BatchNorm3d(oup_inc), nn.ReLU(inplace=True))
            self.banch2 = nn.Sequential(
                nn.Conv3d(inp, oup_inc, 1, 1, 0, bias=False), nn.BatchNorm3d(oup_inc), nn.ReLU(inplace=True),
                nn.Conv3d(oup_inc, oup_inc, 3, (1, stride, stride), 1, groups=oup_inc, bias=False),
                nn.BatchNorm3d(oup_inc),
                nn.Conv3d(oup_inc, oup_inc, 1, 1, 0, bias=False), nn.BatchNorm3d(oup_inc), nn.ReLU(inplace=True))


class ShuffleNetV2_Inverted_Residual_Block(_Holder):
    """shufflenetv2_helper.py:178-219."""

    def __init__(self, input_channel, idxstage, stage_out_channels):
        super().__init__()
        self.stage_repeats = [4, 8, 4]
        feats = []
        output_channel = stage_out_channels[idxstage + 2]
        for i in range(self.stage_repeats[idxstage]):
            feats.append(InvertedResidual(input_channel, output_channel, 2 if i == 0 else 1))
            input_channel = output_channel
        self.features = nn.Sequential(*feats)
        _stage_init(self)


class ShuffleNetV2_Stage(_Holder):
    """shufflenetv2_helper.py:222-297."""

    def __init__(self, input_channel, idxstage, slow_stage_out_channels, fast_stage_out_channels):
        super().__init__()
        self.idxstage = idxstage
        self.num_pathways = len(input_channel)
        self.out_channels = [slow_stage_out_channels[idxstage + 2], fast_stage_out_channels[idxstage + 2]]
        for p, chans in enumerate((slow_stage_out_channels, fast_stage_out_channels)):
            blk = ShuffleNetV2_Inverted_Residual_Block(input_channel[p], idxstage, chans)
            self.add_module("pathway{}_channel_{}".format(p, chans[idxstage + 2]), blk)
            _stage_init(self)

    def pathway(self, p):
        return getattr(self, "pathway{}_channel_{}".format(p, self.out_channels[p]))


class ShuffleNetV2_Model_Stem(_Holder):
    """stem_helper.py:237-270: Conv3d 3x3x3 s(1,2,2) p1 -> BN -> ReLU -> MaxPool3d k3 s(1,2,2) p1."""

    def __init__(self, input_channels, img_dim=3):
        super().__init__()
        self.num_pathways = len(input_channels)
        for p in range(self.num_pathways):
            stem = nn.Sequential(
                nn.Conv3d(img_dim, input_channels[p], kernel_size=3, stride=(1, 2, 2), padding=(1, 1, 1), bias=False),
                nn.BatchNorm3d(input_channels[p]), nn.ReLU(inplace=True),
                nn.MaxPool3d(kernel_size=3, stride=(1, 2, 2), padding=1))
            self.add_module("pathway{}_stem".format(p), stem)


class ShuffleNetV2BasicHead(_Holder):
    """head_helper.py:499-557: per pathway 1x1x1 conv + BN + ReLU -> global avg pool -> cat -> Dropout -> Linear ->
    softmax -> mean."""

    def __init__(self, input_channel, last_channel, num_classes, dropout_rate, act_func="softmax"):
        super().__init__()
        self.num_pathways = len(input_channel)
        for p in range(self.num_pathways):
            feats = nn.Sequential(nn.Sequential(
                nn.Conv3d(input_channel[p], last_channel[p], 1, 1, 0, bias=False), nn.BatchNorm3d(last_channel[p]),
                nn.ReLU(inplace=True)))
            self.add_module("pathway{}_conv1x1x1".format(p), feats)
        if act_func == "softmax":
            self.act = nn.Softmax(dim=4)
        elif act_func == "sigmoid":
            self.act = nn.Sigmoid()
        self.act_func = act_func
        self.classifier = nn.Sequential(nn.Dropout(dropout_rate), nn.Linear(sum(last_channel), num_classes, bias=True))


@MODEL_REGISTRY.register()
class SlowFastShuffleNetV2(_EfficientBase):
    """custom_video_model_builder.py:448-617."""

    def __init__(self, cfg):
        super().__init__()
        self.norm_module = get_norm(cfg)
        if cfg.DETECTION.ENABLE:
            raise NotImplementedError("DETECTION.ENABLE is out of scope")
        self.enable_detection = False
        width_mult = cfg.SLOWFAST.WIDTH_MULTI
        table = {0.25: [-1, 24, 32, 64, 128, 1024], 0.5: [-1, 24, 48, 96, 192, 1024],
                 1.0: [-1, 24, 116, 240, 464, 1024],   # 232 -> 240 in the reference (:476)
                 1.5: [-1, 24, 176, 352, 704, 1024], 2.0: [-1, 24, 224, 496, 976, 2048]}  # 488 -> 496 (:480)
        if width_mult not in table:
            raise ValueError("{} groups is not supported for 1x1 Grouped Convolutions".format(width_mult))
        self.stage_out_channels = table[width_mult]
        beta = cfg.SLOWFAST.BETA_INV
        self.fast_stage_out_channels = [c // beta for c in self.stage_out_channels]
        so, fo = self.stage_out_channels, self.fast_stage_out_channels
        self.s1 = ShuffleNetV2_Model_Stem(input_channels=[so[1], so[1] // beta], img_dim=len(cfg.DATA.MEAN))
        for i in range(1, 5):
            fuse = FuseFastAndSlow(dim_in=[so[i], fo[i]], alpha=cfg.SLOWFAST.ALPHA, beta_inv=beta,
                                   norm_module=self.norm_module)
            setattr(self, "s%d_fuse" % i, fuse)
            if i < 4:
                stage = ShuffleNetV2_Stage(input_channel=[so[i] + fo[i], fo[i] + so[i] // beta], idxstage=i - 1,
                                           slow_stage_out_channels=so, fast_stage_out_channels=fo)
                setattr(self, "s%d" % (i + 1), stage)
        self.head = ShuffleNetV2BasicHead(
            input_channel=[so[4] + fo[4], fo[4] + so[4] // beta], last_channel=[so[-1], fo[-1]],
            num_classes=cfg.MODEL.NUM_CLASSES, dropout_rate=cfg.MODEL.DROPOUT_RATE, act_func=cfg.MODEL.HEAD_ACT)
        init_weights(self, cfg.MODEL.FC_INIT_STD, cfg.RESNET.ZERO_INIT_FINAL_BN)
        self._init_runtime(cfg)

    # child order of the reference: s1, s1_fuse, s2, s2_fuse, s3, s3_fuse, s4, s4_fuse, head -- setattr order above
    # registers s1_fuse before s2 etc., matching it.

    def _emit_unit(self, plan, unit, x, y):
        """InvertedResidual.forward (shufflenetv2_helper.py:104-112)."""
        if unit.stride == 1:
            c = x.shape[4] // 2
            x1, x2 = x[..., :c], x[..., c:]
            b2 = unit.banch2
            t = self._emit_conv(plan, x2, b2[0], b2[1], rt.ACT_RELU)
            t = self._emit_conv(plan, t, b2[3], b2[4], rt.ACT_NONE)
            t = self._emit_conv(plan, t, b2[5], b2[6], rt.ACT_RELU)
            plan.shuffle_concat(x1, t, 2, y)
        else:
            b1, b2 = unit.banch1, unit.banch2
            u = self._emit_conv(plan, x, b1[0], b1[1], rt.ACT_NONE)
            u = self._emit_conv(plan, u, b1[2], b1[3], rt.ACT_RELU)
            t = self._emit_conv(plan, x, b2[0], b2[1], rt.ACT_RELU)
            t = self._emit_conv(plan, t, b2[3], b2[4], rt.ACT_NONE)
            t = self._emit_conv(plan, t, b2[5], b2[6], rt.ACT_RELU)
            plan.shuffle_concat(u, t, 2, y)

    def _compile(self, plan):
        cfg = self._cfg
        alpha, beta = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV
        so, fo = self.stage_out_channels, self.fast_stage_out_channels
        xs_in = plan.inputs
        B = xs_in[0].shape[0]
        assert xs_in[1].shape[2] == xs_in[0].shape[2] * alpha, "fast pathway must have ALPHA x the frames of the slow one"
        # concat buffers after stage i: slow [x_s (so) | from_fast (fo)], fast [from_slow (so // beta) | x_f (fo)]
        cur = []
        for p in range(2):
            seq = getattr(self.s1, "pathway{}_stem".format(p))
            _, _, T, H, W = xs_in[p].shape
            T, H, W = self._pooled_shape(T, H, W, seq)
            c_own = so[1] if p == 0 else fo[1]
            tot = so[1] + fo[1] if p == 0 else so[1] // beta + fo[1]
            off = 0 if p == 0 else so[1] // beta
            buf = plan.act(B, T, H, W, tot, name="s1_cat%d" % p)
            self._emit_stem(plan, seq, xs_in[p], buf[..., off:off + c_own], rt.ACT_RELU)
            cur.append((buf, off, c_own))
        self._emit_fuse(plan, self.s1_fuse, cur)
        for i in range(1, 4):
            stage = getattr(self, "s%d" % (i + 1))
            nxt = []
            for p in range(2):
                x = cur[p][0]
                c_own = so[i + 1] if p == 0 else fo[i + 1]
                tot = so[i + 1] + fo[i + 1] if p == 0 else so[i + 1] // beta + fo[i + 1]
                off = 0 if p == 0 else so[i + 1] // beta
                units = stage.pathway(p).features
                _, T, H, W, _ = x.shape
                Ho, Wo = _conv_out(H, 3, 2, 1), _conv_out(W, 3, 2, 1)
                dst = plan.act(B, T, Ho, Wo, tot, name="s%d_cat%d" % (i + 1, p))
                for ui, unit in enumerate(units):
                    last = ui == len(units) - 1
                    y = dst[..., off:off + c_own] if last else plan.act(B, T, Ho, Wo, c_own)
                    self._emit_unit(plan, unit, x, y)
                    x = y
                nxt.append((dst, off, c_own))
            cur = nxt
            self._emit_fuse(plan, getattr(self, "s%d_fuse" % (i + 1)), cur)
        feats = []
        for p in range(2):
            seq = getattr(self.head, "pathway{}_conv1x1x1".format(p))[0]
            feats.append(self._emit_conv(plan, cur[p][0], seq[0], seq[1], rt.ACT_RELU))
        act = {"softmax": rt.HEAD_SOFTMAX, "sigmoid": rt.HEAD_SIGMOID}[self.head.act_func]
        lin = self.head.classifier[1]
        plan.head(feats, lin.weight, lin.bias, act)


# ================================================================================================ ShuffleNet (v1)
class Bottleneck(_Holder):
    """shufflenet_helper.py:37-84."""

    def __init__(self, in_planes, out_planes, stride, groups):
        super().__init__()
        self.stride = stride
        self.groups = groups
        mid_planes = out_planes // 4
        if self.stride == 2:
            mid_planes = out_planes // 2
            out_planes = out_planes - out_planes // 2
        g = 1 if in_planes == 24 else groups
        self.conv1 = nn.Conv3d(in_planes, mid_planes, kernel_size=1, groups=g, bias=False)
        self.bn1 = nn.BatchNorm3d(mid_planes)
        self.conv2 = nn.Conv3d(mid_planes, mid_planes, kernel_size=(3, 3, 3), stride=(1, stride, stride), padding=1,
                               groups=mid_planes, bias=False)
        self.bn2 = nn.BatchNorm3d(mid_planes)
        self.conv3 = nn.Conv3d(mid_planes, out_planes, kernel_size=1, groups=groups, bias=False)
        self.bn3 = nn.BatchNorm3d(out_planes)
        self.relu = nn.ReLU(inplace=True)
        if stride == 2:
            self.shortcut = nn.Sequential(
                nn.Conv3d(in_planes, mid_planes, kernel_size=1, bias=False),
                nn.AvgPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))


class ShuffleNet_Residual_Block(_Holder):
    """shufflenet_helper.py:165-218."""

    def __init__(self, in_plane, out_plane, num_block, group):
        super().__init__()
        layers = []
        for i in range(num_block):
            layers.append(Bottleneck(in_plane, out_plane, stride=2 if i == 0 else 1, groups=group))
            in_plane = out_plane
        self.features = nn.Sequential(*layers)
        _stage_init(self)


class ShuffleNet_Stage(_Holder):
    """shufflenet_helper.py:221-295."""

    def __init__(self, input_channel, slow_stage_out_channels, fast_stage_out_channels, num_block, group):
        super().__init__()
        self.out_channels = [slow_stage_out_channels, fast_stage_out_channels]
        self.num_pathways = len(input_channel)
        for p in range(self.num_pathways):
            blk = ShuffleNet_Residual_Block(input_channel[p], self.out_channels[p], num_block, group)
            self.add_module("pathway{}_channel_{}".format(p, self.out_channels[p]), blk)
            _stage_init(self)

    def pathway(self, p):
        return getattr(self, "pathway{}_channel_{}".format(p, self.out_channels[p]))


class ShuffleNet_Model_Stem(ShuffleNetV2_Model_Stem):
    """stem_helper.py:274-306 (same layers as the ShuffleNetV2 stem)."""


class ShuffleNetBasicHead(_Holder):
    """head_helper.py:562-609: global avg pool per pathway -> cat -> Dropout -> Linear -> softmax -> mean."""

    def __init__(self, input_channel, num_classes, dropout_rate, act_func="softmax"):
        super().__init__()
        self.num_pathways = len(input_channel)
        if act_func == "softmax":
            self.act = nn.Softmax(dim=4)
        elif act_func == "sigmoid":
            self.act = nn.Sigmoid()
        self.act_func = act_func
        self.classifier = nn.Sequential(nn.Dropout(dropout_rate), nn.Linear(sum(input_channel), num_classes, bias=True))


@MODEL_REGISTRY.register()
class SlowFastShuffleNet(_EfficientBase):
    """custom_video_model_builder.py:620-789."""

    def __init__(self, cfg):
        super().__init__()
        self.norm_module = get_norm(cfg)
        if cfg.DETECTION.ENABLE:
            raise NotImplementedError("DETECTION.ENABLE is out of scope")
        self.enable_detection = False
        groups = cfg.SLOWFAST.GROUPS
        self.num_blocks = [4, 8, 4]
        self.groups = groups
        table = {1: [24, 144, 288, 567], 2: [24, 200, 400, 800], 3: [24, 240, 480, 960], 4: [24, 272, 544, 1088],
                 8: [24, 384, 768, 1536]}
        if groups not in table:
            raise ValueError("{} groups is not supported for 1x1 Grouped Convolutions".format(groups))
        beta = cfg.SLOWFAST.BETA_INV
        self.stage_out_channels = [int(i * cfg.SLOWFAST.WIDTH_MULTI) for i in table[groups]]
        self.fast_stage_out_channels = [c // beta for c in self.stage_out_channels]
        so, fo = self.stage_out_channels, self.fast_stage_out_channels
        self.s1 = ShuffleNet_Model_Stem(input_channels=[so[0], fo[0]], img_dim=len(cfg.DATA.MEAN))
        for i in range(4):
            fuse = FuseFastAndSlow(dim_in=[so[i], fo[i]], alpha=cfg.SLOWFAST.ALPHA, beta_inv=beta,
                                   norm_module=self.norm_module)
            setattr(self, "s%d_fuse" % (i + 1), fuse)
            if i < 3:
                stage = ShuffleNet_Stage(input_channel=[so[i] + fo[i], fo[i] + so[i] // beta],
                                         slow_stage_out_channels=so[i + 1], fast_stage_out_channels=fo[i + 1],
                                         num_block=self.num_blocks[i], group=groups)
                setattr(self, "s%d" % (i + 2), stage)
        self.head = ShuffleNetBasicHead(input_channel=[so[3] + fo[3], fo[3] + so[3] // beta],
                                        num_classes=cfg.MODEL.NUM_CLASSES, dropout_rate=cfg.MODEL.DROPOUT_RATE,
                                        act_func=cfg.MODEL.HEAD_ACT)
        init_weights(self, cfg.MODEL.FC_INIT_STD, cfg.RESNET.ZERO_INIT_FINAL_BN)
        self._init_runtime(cfg)

    def _emit_unit(self, plan, blk, x, y):
        """Bottleneck.forward (shufflenet_helper.py:75-84)."""
        t = self._emit_conv(plan, x, blk.conv1, blk.bn1, rt.ACT_RELU)
        if blk.groups > 1:
            ts = plan.act(*t.shape)
            plan.shuffle_concat(t, None, blk.groups, ts)
            t = ts
        t = self._emit_conv(plan, t, blk.conv2, blk.bn2, rt.ACT_NONE)
        if blk.stride == 2:
            c3 = blk.conv3.out_channels
            self._emit_conv(plan, t, blk.conv3, blk.bn3, rt.ACT_RELU, y=y[..., :c3])      # relu(cat) == cat(relu)
            sc = self._emit_conv(plan, x, blk.shortcut[0], None, rt.ACT_NONE)
            plan.pool(sc, y[..., c3:], (1, 3, 3), (1, 2, 2), (0, 1, 1), is_avg=True, act=rt.ACT_RELU)
        else:
            self._emit_conv(plan, t, blk.conv3, blk.bn3, rt.ACT_RELU, y=y, res=x)

    def _compile(self, plan):
        cfg = self._cfg
        alpha, beta = cfg.SLOWFAST.ALPHA, cfg.SLOWFAST.BETA_INV
        so, fo = self.stage_out_channels, self.fast_stage_out_channels
        xs_in = plan.inputs
        B = xs_in[0].shape[0]
        assert xs_in[1].shape[2] == xs_in[0].shape[2] * alpha, "fast pathway must have ALPHA x the frames of the slow one"
        cur = []
        for p in range(2):
            seq = getattr(self.s1, "pathway{}_stem".format(p))
            _, _, T, H, W = xs_in[p].shape
            T, H, W = self._pooled_shape(T, H, W, seq)
            c_own = so[0] if p == 0 else fo[0]
            tot = so[0] + fo[0] if p == 0 else so[0] // beta + fo[0]
            off = 0 if p == 0 else so[0] // beta
            buf = plan.act(B, T, H, W, tot, name="s1_cat%d" % p)
            self._emit_stem(plan, seq, xs_in[p], buf[..., off:off + c_own], rt.ACT_RELU)
            cur.append((buf, off, c_own))
        self._emit_fuse(plan, self.s1_fuse, cur)
        for i in range(3):
            stage = getattr(self, "s%d" % (i + 2))
            nxt = []
            for p in range(2):
                x = cur[p][0]
                c_own = so[i + 1] if p == 0 else fo[i + 1]
                tot = so[i + 1] + fo[i + 1] if p == 0 else so[i + 1] // beta + fo[i + 1]
                off = 0 if p == 0 else so[i + 1] // beta
                units = stage.pathway(p).features
                _, T, H, W, _ = x.shape
                Ho, Wo = _conv_out(H, 3, 2, 1), _conv_out(W, 3, 2, 1)
                dst = plan.act(B, T, Ho, Wo, tot, name="s%d_cat%d" % (i + 2, p))
                for ui, unit in enumerate(units):
                    last = ui == len(units) - 1
                    y = dst[..., off:off + c_own] if last else plan.act(B, T, Ho, Wo, c_own)
                    self._emit_unit(plan, unit, x, y)
                    x = y
                nxt.append((dst, off, c_own))
            cur = nxt
            self._emit_fuse(plan, getattr(self, "s%d_fuse" % (i + 2)), cur)
        act = {"softmax": rt.HEAD_SOFTMAX, "sigmoid": rt.HEAD_SIGMOID}[self.head.act_func]
        lin = self.head.classifier[1]
        plan.head([c[0] for c in cur], lin.weight, lin.bias, act)
