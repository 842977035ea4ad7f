"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the batched clip forward path.

A functional, state_dict-driven restatement (torch CPU ops, FP32 or FP64) of the
reference's eval-mode forward for the ResNet based models:

  * SlowFastDualAttention  (SlowFast/slowfast/models/custom_video_model_builder.py:171-445)
  * SlowFast               (SlowFast/slowfast/models/video_model_builder.py:153-416)
  * ResNet                 (SlowFast/slowfast/models/video_model_builder.py:419-611; C2D / I3D / Slow, with Nonlocal blocks)

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this file.  The product package never does: the
shipped path is CUDA-only and raises when its extension is missing.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the pin is (a) `tests/test_oracle_vs_reference.py`, which
imports the real reference from /root/reference in the build container and
compares every stage output, and (b) `tests/golden/*.npz`, produced by running
the reference itself (`tests/golden/make_golden.py`), which travel to the GPU box.

All tensors are NCDHW like the reference.  `sd` is a reference-schema
state_dict.  Every function cites the reference lines it
restates.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # every BatchNorm3d on the path is built with eps=1e-5 (resnet_helper.py:124, stem_helper.py:18)

# custom_video_model_builder.py:151-168 / video_model_builder.py:16-90
STAGE_DEPTH = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}   # video_model_builder.py:15-16
TEMPORAL_KERNEL_BASIS = {
    "c2d": [[[1]], [[1]], [[1]], [[1]], [[1]]],
    "c2d_nopool": [[[1]], [[1]], [[1]], [[1]], [[1]]],
    "i3d": [[[5]], [[3]], [[3, 1]], [[3, 1]], [[1, 3]]],
    "i3d_nopool": [[[5]], [[3]], [[3, 1]], [[3, 1]], [[1, 3]]],
    "slow": [[[1]], [[1]], [[1]], [[3]], [[3]]],
    "fast": [[[5]], [[3]], [[3]], [[3]], [[3]]],          # video_model_builder.py:79-85
    "slowfast": [[[1], [5]], [[1], [3]], [[1], [3]], [[3], [3]], [[3], [3]]],
}
POOL1 = {"c2d": [[2, 1, 1]], "c2d_nopool": [[1, 1, 1]], "i3d": [[2, 1, 1]], "i3d_nopool": [[1, 1, 1]],
         "slow": [[1, 1, 1]], "slowfast": [[1, 1, 1], [1, 1, 1]], "fast": [[1, 1, 1]]}


# Where the restatement runs.  "cpu" always -- it is the oracle -- except in tests/experiments/eager_torch_gpu.py, which
# times the same torch ops (cuDNN / cuBLAS eager kernels, what the reference itself would launch) on the GPU.
DEVICE = "cpu"

def _bn(x, sd, p):
    """Eval-mode BatchNorm3d (batchnorm_helper.py:15-24 -> nn.BatchNorm3d)."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    m, v = sd[p + ".running_mean"], sd[p + ".running_var"]
    return F.batch_norm(x, m.to(x.dtype), v.to(x.dtype), w.to(x.dtype), b.to(x.dtype), False, 0.0, BN_EPS)


def _conv(x, sd, p, stride=1, padding=0, dilation=1, groups=1):
    w = sd[p + ".weight"].to(x.dtype)
    b = sd.get(p + ".bias")
    return F.conv3d(x, w, None if b is None else b.to(x.dtype), stride, padding, dilation, groups)


def resnet_stem(x, sd, p):
    """ResNetBasicStem.forward (stem_helper.py:173-178): conv -> BN -> ReLU -> MaxPool(1,3,3)/s(1,2,2)/p(0,1,1).
    Kernel is read from the weight; stride (1,2,2) and padding (k_t//2,3,3) per
    custom_video_model_builder.py:219-228."""
    kt, kh, kw = sd[p + ".conv.weight"].shape[2:]
    x = _conv(x, sd, p + ".conv", stride=(1, 2, 2), padding=(kt // 2, kh // 2, kw // 2))
    x = F.relu(_bn(x, sd, p + ".bn"))
    return F.max_pool3d(x, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))


def bottleneck_block(x, sd, p, stride, dilation=1, num_groups=1, stride_1x1=False):
    """ResBlock.forward (resnet_helper.py:352-358) around BottleneckTransform.forward
    (resnet_helper.py:225-240): [1x1x1 s proj + BN] + (Tx1x1->BN->ReLU->1x3x3 s->BN->ReLU->1x1x1->BN) -> add -> ReLU."""
    s1, s3 = (stride, 1) if stride_1x1 else (1, stride)
    kt = sd[p + ".branch2.a.weight"].shape[2]
    y = _conv(x, sd, p + ".branch2.a", stride=(1, s1, s1), padding=(kt // 2, 0, 0))
    y = F.relu(_bn(y, sd, p + ".branch2.a_bn"))
    y = _conv(y, sd, p + ".branch2.b", stride=(1, s3, s3), padding=(0, dilation, dilation),
              dilation=(1, dilation, dilation), groups=num_groups)
    y = F.relu(_bn(y, sd, p + ".branch2.b_bn"))
    y = _bn(_conv(y, sd, p + ".branch2.c"), sd, p + ".branch2.c_bn")
    if (p + ".branch1.weight") in sd:
        x = _bn(_conv(x, sd, p + ".branch1", stride=(1, stride, stride)), sd, p + ".branch1_bn")
    return F.relu(x + y)


def nonlocal_block(x, sd, p, pool_size, instantiation):
    """Nonlocal.forward (nonlocal_helper.py:105-148): theta / phi / g 1x1x1 convs (phi, g on the max-pooled input),
    affinity einsum("nct,ncp->ntp"), softmax with dim_inner^-0.5 or division by the number of keys, second einsum,
    conv_out -> BN -> + identity."""
    identity = x
    N, C, T, H, W = x.shape
    theta = _conv(x, sd, p + ".conv_theta")
    if pool_size is not None and any(v > 1 for v in pool_size):
        x = F.max_pool3d(x, kernel_size=list(pool_size), stride=list(pool_size), padding=0)
    phi = _conv(x, sd, p + ".conv_phi")
    g = _conv(x, sd, p + ".conv_g")
    d = theta.shape[1]
    theta, phi, g = theta.view(N, d, -1), phi.view(N, d, -1), g.view(N, d, -1)
    theta_phi = torch.einsum("nct,ncp->ntp", (theta, phi))
    if instantiation == "softmax":
        theta_phi = theta_phi * (d ** -0.5)
        theta_phi = F.softmax(theta_phi, dim=2)
    elif instantiation == "dot_product":
        theta_phi = theta_phi / theta_phi.shape[2]
    else:
        raise NotImplementedError("Unknown norm type {}".format(instantiation))
    y = torch.einsum("ntg,ncg->nct", (theta_phi, g)).view(N, d, T, H, W)
    return identity + _bn(_conv(y, sd, p + ".conv_out"), sd, p + ".bn")


def res_stage(xs, sd, p, num_blocks, strides, dilations, num_groups=(1, 1), stride_1x1=False, nonlocal_group=None,
              nonlocal_pool=None, instantiation="softmax"):
    """ResStage.forward (resnet_helper.py:530-561).  A Nonlocal block follows residual block i whenever the weights
    hold `pathway{pw}_nonlocal{i}` (the reference tests hasattr the same way); NONLOCAL.GROUP > 1 folds the temporal
    axis into the batch around it (resnet_helper.py:541-560)."""
    out = []
    for pw, x in enumerate(xs):
        for i in range(num_blocks[pw]):
            x = bottleneck_block(x, sd, "%s.pathway%d_res%d" % (p, pw, i), strides[pw] if i == 0 else 1,
                                 dilations[pw], num_groups[pw], stride_1x1)
            q = "%s.pathway%d_nonlocal%d" % (p, pw, i)
            if (q + ".conv_theta.weight") in sd:
                grp = nonlocal_group[pw] if nonlocal_group is not None else 1
                b, c, t, h, w = x.shape
                if grp > 1:
                    x = x.permute(0, 2, 1, 3, 4).reshape(b * grp, t // grp, c, h, w).permute(0, 2, 1, 3, 4)
                x = nonlocal_block(x, sd, q, nonlocal_pool[pw] if nonlocal_pool is not None else [1, 2, 2], instantiation)
                if grp > 1:
                    x = x.permute(0, 2, 1, 3, 4).reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
        out.append(x)
    return out


def position_attention(x, sd, p, row_chunk=2048):
    """SpatialAttention.forward (wdf_attention_helper.py:33-54), query-row-chunked.

    A = softmax_j(Q^T K) over all N = T*H*W key positions (no 1/sqrt(d) scale),
    O = V A^T, out = gamma * O + x.  Rows of A are independent, so evaluating
    row blocks of `row_chunk` queries is the same arithmetic without the N x N
    matrix (SURVEY.md section 8c 'Large-N oracle')."""
    B, C, T, H, W = x.shape
    N = T * H * W
    q = _conv(x, sd, p + ".query_conv").reshape(B, -1, N)   # (B, dq, N)
    k = _conv(x, sd, p + ".key_conv").reshape(B, -1, N)     # (B, dq, N)
    v = _conv(x, sd, p + ".value_conv").reshape(B, -1, N)   # (B, C,  N)
    out = torch.empty(B, C, N, dtype=x.dtype, device=x.device)
    for n0 in range(0, N, row_chunk):
        n1 = min(N, n0 + row_chunk)
        att = torch.softmax(torch.bmm(q[:, :, n0:n1].transpose(1, 2), k), dim=-1)  # (B, n, N)
        out[:, :, n0:n1] = torch.bmm(v, att.transpose(1, 2))
    gamma = sd[p + ".gamma"].to(x.dtype)
    return gamma * out.reshape(B, C, T, H, W) + x


def eca(x, sd, p):
    """ECA.forward (wdf_attention_helper.py:77-91): GAP -> Conv1d(1,1,k,pad k//2, no bias) across the channel
    axis (zero padded) -> sigmoid -> channel scale."""
    w = sd[p + ".conv.weight"].to(x.dtype)          # (1,1,k)
    y = x.mean(dim=(2, 3, 4))                       # (B, C)
    y = F.conv1d(y.unsqueeze(1), w, padding=(w.shape[-1] - 1) // 2).squeeze(1)
    return x * torch.sigmoid(y)[:, :, None, None, None]


def fuse_fast_and_slow(xs, sd, p, alpha):
    """FuseFastAndSlow.forward = CMDA (custom_video_model_builder.py:123-148).
    fast->slow: MaxPool(alpha,1,1) -> ECA -> BN -> ReLU -> cat([x_s, .]);
    slow->fast: 1x1x1 conv C->C/beta -> position attention -> BN -> ReLU -> nearest x alpha in T -> cat([., x_f])."""
    x_s, x_f = xs
    f = F.max_pool3d(x_f, kernel_size=(alpha, 1, 1), stride=(alpha, 1, 1))
    f = eca(f, sd, p + ".attention_channel_f2s")
    f = F.relu(_bn(f, sd, p + ".bn_f2s"))
    s = _conv(x_s, sd, p + ".downsample_c_of_slow")
    s = position_attention(s, sd, p + ".attention_spatial_s2f")
    s = F.relu(_bn(s, sd, p + ".bn_s2f"))
    s = s.repeat_interleave(alpha, dim=2)           # nn.Upsample(scale=(alpha,1,1), 'nearest')
    return [torch.cat([x_s, f], 1), torch.cat([s, x_f], 1)]


def fuse_fast_to_slow(xs, sd, p, alpha):
    """FuseFastToSlow.forward (video_model_builder.py:143-150): conv kx1x1 stride (alpha,1,1) pad k//2 -> BN -> ReLU -> cat."""
    x_s, x_f = xs
    k = sd[p + ".conv_f2s.weight"].shape[2]
    f = _conv(x_f, sd, p + ".conv_f2s", stride=(alpha, 1, 1), padding=(k // 2, 0, 0))
    f = F.relu(_bn(f, sd, p + ".bn"))
    return [torch.cat([x_s, f], 1), x_f]


def resnet_basic_head(xs, sd, p, pool_sizes, act="softmax", return_logits=False):
    """ResNetBasicHead.forward eval branch (head_helper.py:198-223): per-pathway AvgPool3d(pool, stride 1)
    (or global when pool is None) -> cat -> NTHWC -> Linear -> softmax(dim=4) -> mean(T,H,W)."""
    pooled = []
    for x, ps in zip(xs, pool_sizes):
        pooled.append(F.adaptive_avg_pool3d(x, 1) if ps is None else F.avg_pool3d(x, ps, stride=1))
    x = torch.cat(pooled, 1).permute(0, 2, 3, 4, 1)
    logits = F.linear(x, sd[p + ".projection.weight"].to(x.dtype), sd[p + ".projection.bias"].to(x.dtype))
    if act == "softmax":
        y = torch.softmax(logits, dim=4)
    elif act == "sigmoid":
        y = torch.sigmoid(logits)
    else:
        raise NotImplementedError(act)
    y = y.mean([1, 2, 3]).reshape(x.shape[0], -1)
    if return_logits:
        return y, logits
    return y


def _head_pools(cfg):
    """Head pool kernels (custom_video_model_builder.py:404-424 / video_model_builder.py:375-394)."""
    if cfg.MULTIGRID.SHORT_CYCLE:
        return [None, None]
    ps = POOL1[cfg.MODEL.ARCH]
    c = cfg.DATA.CROP_SIZE // 32
    return [[cfg.DATA.NUM_FRAMES // cfg.SLOWFAST.ALPHA // ps[0][0], c // ps[0][1], c // ps[0][2]],
            [cfg.DATA.NUM_FRAMES // ps[1][0], c // ps[1][1], c // ps[1][2]]]


def _resnet_two_stream(cfg, sd, inputs, fuse, dtype, taps):
    assert len(inputs) == 2, "Input tensor does not contain 2 pathway"
    depth = STAGE_DEPTH[cfg.RESNET.DEPTH]
    ng = [cfg.RESNET.NUM_GROUPS] * 2
    sd = {k: v.detach().to(DEVICE) for k, v in sd.items()}
    xs = [t.detach().to(DEVICE, dtype) for t in inputs]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val] if isinstance(val, list) else val.clone()

    xs = [resnet_stem(xs[0], sd, "s1.pathway0_stem"), resnet_stem(xs[1], sd, "s1.pathway1_stem")]
    tap("s1", xs)
    xs = fuse(xs, sd, "s1_fuse")
    tap("s1_fuse", xs)
    for i, stage in enumerate(("s2", "s3", "s4", "s5")):
        # pathway{0,1}_pool after s2_fuse is MaxPool3d k=s=[1,1,1] == identity for arch 'slowfast'
        # (custom_video_model_builder.py:278-284, _POOL1 :165-167)
        xs = res_stage(xs, sd, stage, [depth[i]] * 2, cfg.RESNET.SPATIAL_STRIDES[i],
                       cfg.RESNET.SPATIAL_DILATIONS[i], ng, cfg.RESNET.STRIDE_1X1)
        tap(stage, xs)
        if stage != "s5":
            xs = fuse(xs, sd, stage + "_fuse")
            tap(stage + "_fuse", xs)
    y, logits = resnet_basic_head(xs, sd, "head", _head_pools(cfg), cfg.MODEL.HEAD_ACT, return_logits=True)
    tap("logits", logits)
    tap("head", y)
    return y


def slowfast_dual_attention_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFastDualAttention.forward (custom_video_model_builder.py:428-445), eval mode."""
    a = cfg.SLOWFAST.ALPHA
    return _resnet_two_stream(cfg, sd, inputs, lambda xs, s, p: fuse_fast_and_slow(xs, s, p, a), dtype, taps)


def slowfast_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """SlowFast.forward (video_model_builder.py:399-416), eval mode."""
    a = cfg.SLOWFAST.ALPHA
    return _resnet_two_stream(cfg, sd, inputs, lambda xs, s, p: fuse_fast_to_slow(xs, s, p, a), dtype, taps)


def resnet_forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """ResNet.forward (video_model_builder.py:599-611): single pathway C2D / I3D / Slow (+ _NLN), eval mode.
    The temporal kernels of the blocks are read from the weights; the only arch-dependent step is the max-pool after
    res2 (kernel = stride = _POOL1[arch], video_model_builder.py:503-509)."""
    assert len(inputs) == 1, "Input tensor does not contain 1 pathway"
    depth = STAGE_DEPTH[cfg.RESNET.DEPTH]
    sd = {k: v.detach().to(DEVICE) for k, v in sd.items()}
    xs = [inputs[0].detach().to(DEVICE, dtype)]

    def tap(name, val):
        if taps is not None:
            taps[name] = [t.clone() for t in val]

    xs = [resnet_stem(xs[0], sd, "s1.pathway0_stem")]
    tap("s1", xs)
    pool = POOL1[cfg.MODEL.ARCH][0]
    for i, stage in enumerate(("s2", "s3", "s4", "s5")):
        xs = res_stage(xs, sd, stage, [depth[i]], cfg.RESNET.SPATIAL_STRIDES[i], cfg.RESNET.SPATIAL_DILATIONS[i],
                       [cfg.RESNET.NUM_GROUPS], cfg.RESNET.STRIDE_1X1, cfg.NONLOCAL.GROUP[i], cfg.NONLOCAL.POOL[i],
                       cfg.NONLOCAL.INSTANTIATION)
        tap(stage, xs)
        if stage == "s2":
            xs = [F.max_pool3d(xs[0], kernel_size=pool, stride=pool, padding=0)]
    assert not cfg.MULTIGRID.SHORT_CYCLE, "pathway dimensions are not consistent."   # as the reference head asserts
    c = cfg.DATA.CROP_SIZE // 32
    pools = [[cfg.DATA.NUM_FRAMES // pool[0], c // pool[1], c // pool[2]]]
    y, logits = resnet_basic_head(xs, sd, "head", pools, cfg.MODEL.HEAD_ACT, return_logits=True)
    if taps is not None:
        taps["logits"], taps["head"] = logits.clone(), y.clone()
    return y


FORWARDS = {
    "SlowFastDualAttention": slowfast_dual_attention_forward,
    "SlowFast": slowfast_forward,
    "ResNet": resnet_forward,
}


def forward(cfg, sd, inputs, dtype=torch.float32, taps=None):
    """Dispatch on cfg.MODEL.MODEL_NAME over every model the oracle restates (R50 models here, efficient models in
    efficient_oracle.py)."""
    fwd = FORWARDS.get(cfg.MODEL.MODEL_NAME)
    if fwd is None:
        from . import efficient_oracle
        fwd = efficient_oracle.FORWARDS[cfg.MODEL.MODEL_NAME]
    with torch.no_grad():
        return fwd(cfg, sd, inputs, dtype=dtype, taps=taps)


def pack_pathway_output(frames, alpha):
    """datasets/utils.py:73-112 for arch 'slowfast': fast = frames, slow = index_select(linspace(0,T-1,T//alpha).long())
    along the time axis.  frames: (..., C, T, H, W) with T at dim -3."""
    T = frames.shape[-3]
    idx = torch.linspace(0, T - 1, T // alpha).long()
    return [frames.index_select(frames.dim() - 3, idx), frames]
