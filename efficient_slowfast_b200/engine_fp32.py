"""FP32-accurate launch plan (cfg.ESF.PRECISION = "fp32"): the north star's "FP32/TF32 path, rel err <= 1e-4".

The reference computes in FP32 everywhere (SlowFast/slowfast/models/resnet_helper.py:182-240, wdf_attention_helper.py:
42-53).  The tensor cores take 16-bit operands, so this plan carries every GEMM operand as a PAIR of FP16 numbers
x = hi + lo (22 mantissa bits) and evaluates x.w = x_hi.w_hi + x_lo.w_hi + x_hi.w_lo with FP32 accumulation -- on the
SAME tcgen05 implicit-GEMM kernel as the 16-bit plan, with no kernel change:

  * an activation is an `Act32`: the FP32 tensor (what the element-wise kernels read: residuals, pooling, ECA, head)
    plus a channels-last FP16 buffer of three planes [hi | lo | hi] (what the GEMM reads as 3 C input channels);
  * a folded weight becomes [w_hi | w_hi | w_lo] along its input-channel axis, rows pre-scaled by a power of two so that
    w_lo is a normal FP16 number;
  * every conv is `igemm (FP32 out, raw accumulators) -> esf_p32_post (un-scale + bias + residual + activation -> FP32
    tensor + the three planes)`.
Concat stays free: a producer writes its channel slice of the FP32 buffer and of each plane.  The position attention
runs in FP32 on the CUDA cores (esf_p32_attention, flash style): the tcgen05 attention kernel rounds P and V to FP16,
which alone costs 5e-4 on the probabilities (measured).

Scope: the two ResNet-50 two-stream models and the single-pathway ResNet (C2D / I3D / Slow), Non-local blocks included.  About 3x the
tensor-core work and ~5x the activation bytes of the FP16 plan: an accuracy mode, benchmarked beside it.
"""
import ctypes
import os

import torch

from . import runtime as rt
from .engine import Plan, bn_affine, fold_conv_bn, host64


class Act32:
    """One activation of the FP32-accurate plan: `f32` (B,T,H,W,C) FP32 view and the [hi|lo|hi] FP16 buffer `x3`
    (B,T,H,W,3*plane) it is mirrored into at channel offset `c0` of every plane.  Supports the `[..., a:b]` channel
    slicing the model code uses for concat buffers."""

    def __init__(self, f32, x3, plane, c0, c_total, weight_order=False, has_f32=True, has_planes=True):
        self.f32, self.x3, self.plane, self.c0, self.c_total = f32, x3, plane, c0, c_total
        self.weight_order = weight_order    # planes [hi | hi | lo]: the tensor is the WEIGHT operand of a per-clip GEMM
        # an intermediate that is only ever a GEMM operand keeps no FP32 copy (f32 is a shape-only meta tensor), one that
        # is only ever a residual keeps no planes (PrecisePlan.act(role=...)): 4 resp. 6 bytes per element the post-pass
        # does not write
        self.has_f32, self.has_planes = has_f32, has_planes

    @property
    def shape(self):
        return self.f32.shape

    @property
    def whole(self):
        return self.c0 == 0 and self.f32.shape[4] == self.c_total

    @property
    def hi(self):
        return self.x3[..., self.c0:self.c0 + self.f32.shape[4]]

    def __getitem__(self, idx):
        assert isinstance(idx, tuple) and len(idx) == 2 and idx[0] is Ellipsis and isinstance(idx[1], slice)
        start, stop, step = idx[1].indices(self.f32.shape[4])
        assert step == 1
        return Act32(self.f32[..., start:stop], self.x3, self.plane, self.c0 + start, self.c_total, self.weight_order,
                     self.has_f32, self.has_planes)

    def need_f32(self, who):
        assert self.has_f32, "%s reads the FP32 copy of an activation that was planned without one" % who
        return self.f32


def split_weight_rows(w):
    """Folded FP64 weight (Cout, Cin, kT, kH, kW) -> (w_hi, w_lo, inv_scale): rows scaled by 2^k so that max|row| lies in
    [1024, 2048) -- w_lo = fp16(w s - w_hi) ~ 2^-12 of that stays far above the FP16 subnormals -- then split into two
    FP16-representable parts; inv_scale[n] = 2^-k undoes the scaling in the FP32 post-pass (exact)."""
    cout = w.shape[0]
    amax = w.reshape(cout, -1).abs().amax(dim=1).clamp_min(1e-30)
    k = torch.floor(torch.log2(1024.0 / amax))
    s = torch.pow(torch.tensor(2.0, dtype=torch.float64), k)
    ws = w * s.view(-1, 1, 1, 1, 1)
    hi = ws.to(torch.float16).to(torch.float64)
    lo = (ws - hi).to(torch.float16).to(torch.float64)
    return hi, lo, 1.0 / s


class PrecisePlan(Plan):
    def __init__(self, device, precision="fp32"):
        super().__init__(device, precision="fp32")
        self.wfold = False           # the banded thin-layer GEMM has no FP32 output; every conv is the plain igemm
        self.precise = True

    # ---------------------------------------------------------------- memory
    def act(self, B, T, H, W, C, name=None, dtype=None, role=None):
        """role "operand": the tensor is only read as a GEMM operand (planes, no FP32 copy); "residual": only as a
        residual (FP32, no planes); None: both."""
        if dtype is not None and dtype != self.adt:
            return super().act(B, T, H, W, C, name=name, dtype=dtype)
        assert role in (None, "operand", "residual") and not (name and role)
        plane = (C + 7) // 8 * 8
        has_f32, has_planes = role != "operand", role != "residual"
        f32 = torch.empty((B, T, H, W, C), dtype=torch.float32, device=self.device if has_f32 else "meta")
        # zero-filled: the row padding [C, plane) of every plane meets zero weights in the GEMM and must stay finite
        x3 = torch.zeros((B, T, H, W, 3 * plane), dtype=torch.float16, device=self.device) if has_planes else None
        self.keep += [f32, x3]
        if name:
            self.buffers[name] = f32
        return Act32(f32, x3, plane, 0, C, has_f32=has_f32, has_planes=has_planes)

    def _post(self, acc, y, scale=None, bias=None, res=None, act=rt.ACT_NONE, label=""):
        """esf_p32_post: y.f32 = act(acc * scale + bias + res) and its three FP16 planes (y: Act32 or FP32 tensor)."""
        L = rt.lib()
        av = rt.view(acc)
        rv = rt.view(res.need_f32("a residual add") if isinstance(res, Act32) else res) if res is not None else rt.null_view()
        if isinstance(y, Act32):
            in_place = y.has_f32 and y.f32.data_ptr() == acc.data_ptr() and scale is None and bias is None and \
                res is None and act == rt.ACT_NONE
            y32 = rt.null_view() if (in_place or not y.has_f32) else rt.view(y.f32)
            y3, plane = (rt.view(y.hi) if y.has_planes else rt.null_view()), y.plane
            worder = int(y.weight_order)
        else:
            y32, y3, plane, worder = rt.view(y), rt.null_view(), 0, 0
        sc = scale.data_ptr() if scale is not None else None
        bs = bias.data_ptr() if bias is not None else None
        self.keep += [av, rv, y32, y3, scale, bias]
        n = acc.numel()
        self._add(lambda s: rt.check(L.esf_p32_post(ctypes.byref(av), sc, bs, ctypes.byref(rv), act, ctypes.byref(y32),
                                                    ctypes.byref(y3), plane, worder, s), "esf_p32_post"),
                  "p32_post", label, nbytes=n * (4 + (4 if res is not None else 0) + (4 if y32.ptr else 0) +
                                                  (6 if y3.ptr else 0)))

    # ---------------------------------------------------------------- ops
    def conv(self, x, y, w_folded, bias, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1), groups=1,
             act=rt.ACT_NONE, res=None, out_dtype=None):
        """Conv3d + folded BN (+ residual) + activation at ~FP32 accuracy: igemm over the [hi|lo|hi] planes with the
        [w_hi|w_hi|w_lo] weight (raw FP32 accumulators), then the FP32 post-pass."""
        if groups != 1:
            raise NotImplementedError("grouped convolutions are not part of the FP32-accurate plan (R50 models only)")
        assert isinstance(x, Act32) and not x.weight_order
        assert x.has_planes, "a convolution reads the planes of an activation that was planned without them"
        w = w_folded.to(torch.float64)
        cout, cin = w.shape[:2]
        assert cin == x.shape[4]
        hi, lo, inv_scale = split_weight_rows(w)
        # the GEMM reads the WHOLE parent buffer: a channel slice (x_s of a concat buffer) gets zero weights elsewhere
        plane = x.plane
        w3 = torch.zeros((cout, 3 * plane) + tuple(w.shape[2:]), dtype=torch.float64)
        w3[:, x.c0:x.c0 + cin] = hi
        w3[:, plane + x.c0:plane + x.c0 + cin] = hi
        w3[:, 2 * plane + x.c0:2 * plane + x.c0 + cin] = lo
        yshape = tuple(y.shape)
        acc = self.scratch(yshape, torch.float32)
        zero = torch.zeros(cout, dtype=torch.float64)
        self.conv_igemm(x.x3, acc, w3, zero, stride, padding, dilation, rt.ACT_NONE, None)
        self.meta[-1]["label"] += " x3"
        self._post(acc, y, self.tensor(inv_scale), self.tensor(bias), res, act)

    def stem(self, x_nc, y, w_folded, bias, stride, padding, act=rt.ACT_RELU):
        """Stem conv + folded BN + ReLU at FP32 accuracy.  Banded tensor-core GEMM (Plan.stem) when the geometry
        allows: the clip is packed twice (hi and lo halves), the band weight is split like any other weight, and the
        three products hi.hi, lo.hi, hi.lo are three launches of the same GEMM with FP32 outputs (its tap table cannot
        hold 3 x kT x kH taps), summed, un-scaled, biased and activated by esf_p32_post3.  Else the FP32 CUDA-core stem."""
        from .engine import pack_stem_band, pack_stem_tband
        B, Cin, T, H, W = x_nc.shape
        cout = w_folded.shape[0]
        kt, kh, kw = w_folded.shape[2:]
        geo = rt.stem_geometry(W, Cin, kw, stride[2], padding[2]) if stride[0] == 1 else None
        if geo is None or not y.f32.is_contiguous():
            self.stem_conv(x_nc, y.f32, w_folded, bias, stride, padding, act)
            return self._post(y.f32, y, label="split")
        pitch, lpad, _ = geo
        L = rt.lib()
        hi, lo, inv_scale = split_weight_rows(w_folded.to(torch.float64))
        zero = torch.zeros(cout, dtype=torch.float64)
        xps, accs = [], []
        for part in range(2):
            xp = torch.empty((B, T, H, pitch), dtype=torch.float16, device=self.device)
            self.keep.append(xp)
            fn = L.esf_stem_pack_lo if part else L.esf_stem_pack
            self._add(lambda s, xp=xp, fn=fn: rt.check(fn(self._in_ptr(x_nc), B, Cin, T, H, W, pitch, lpad, rt.F16,
                                                          xp.data_ptr(), s), "esf_stem_pack"),
                      "stem_pack", "lo" if part else "hi", nbytes=self._nbytes(x_nc, xp), eager=True)
            xps.append(xp)
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        # kT > 1: the temporal-band kernel (Plan.stem), one launch per split product as well
        twb = rt.stem_tband_wb(W, Cin, cout, kt, kh, kw, stride[2], padding[2]) \
            if os.environ.get("ESF_STEM_TBAND", "1") != "0" else 0
        for name, xp, wpart in (("hi.hi", xps[0], hi), ("lo.hi", xps[1], hi), ("hi.lo", xps[0], lo)):
            if twb:
                wb, bt = pack_stem_tband(wpart, zero, twb, stride[2], self.device, torch.float16)
            else:
                wb, bt = pack_stem_band(wpart, zero, stride[2], self.device, torch.float16)
            create = L.esf_stem_tband_create if twb else L.esf_stem_igemm_create
            acc = torch.empty(tuple(y.shape), dtype=torch.float32, device=self.device)
            self.keep += [wb, bt, acc]
            yv = rt.view(acc)
            h = ctypes.c_void_p()
            rt.check(create(xp.data_ptr(), B, Cin, T, H, W, pitch, wb.data_ptr(), bt.data_ptr(), cout,
                            kt, kh, kw, stride[1], stride[2], padding[0], padding[1], padding[2],
                            rt.ACT_NONE, ctypes.byref(yv), ctypes.byref(h)), "stem GEMM create")
            self.handles.append(h)
            self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "stem_igemm",
                      "%dx%dx%d %d->%d %s %s" % (kt, kh, kw, Cin, cout, "t-band" if twb else "banded", name),
                      flops=2.0 * m * cout * Cin * kt * kh * kw, nbytes=self._nbytes(xp, acc) + wb.numel() * 2)
            accs.append(acc)
        sc, bs = self.tensor(inv_scale), self.tensor(bias)
        avs = [rt.view(a) for a in accs]
        y32, y3 = rt.view(y.f32), rt.view(y.hi)
        self.keep += avs + [y32, y3]
        self._add(lambda s: rt.check(L.esf_p32_post3(ctypes.byref(avs[0]), ctypes.byref(avs[1]), ctypes.byref(avs[2]),
                                                     sc.data_ptr(), bs.data_ptr(), act, ctypes.byref(y32),
                                                     ctypes.byref(y3), y.plane, s), "esf_p32_post3"),
                  "p32_post", "stem sum", nbytes=accs[0].numel() * (12 + 4 + 6))

    def pool(self, x, y, kernel, stride, padding, is_avg=False, act=rt.ACT_NONE):
        assert act == rt.ACT_NONE
        xv, yv = rt.view(x.f32), rt.view(y.f32)
        self.keep += [xv, yv]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_p32_pool3d(ctypes.byref(xv), ctypes.byref(yv), *kernel, *stride, *padding,
                                                      int(is_avg), s), "esf_p32_pool3d"),
                  "p32_pool", "%s C=%d" % (tuple(kernel), x.shape[4]), nbytes=self._nbytes(x.f32, y.f32))
        self._post(y.f32, y, label="split")

    def eca_fuse(self, x_fast, y_slice, alpha, eca_weight, bn):
        scale, shift = bn_affine(bn)
        w = self.tensor(host64(eca_weight).reshape(-1))
        sc, sh = self.tensor(scale), self.tensor(shift)
        B, C = x_fast.shape[0], x_fast.shape[4]
        L = rt.lib()
        scratch = torch.empty(int(L.esf_p32_eca_scratch_floats(B, C)), dtype=torch.float32, device=self.device)
        xv, yv = rt.view(x_fast.f32), rt.view(y_slice.f32)
        self.keep += [scratch, xv, yv]
        k = int(w.numel())
        self._add(lambda s: rt.check(
            L.esf_p32_eca_fuse(ctypes.byref(xv), alpha, w.data_ptr(), k, sc.data_ptr(), sh.data_ptr(),
                               scratch.data_ptr(), ctypes.byref(yv), s), "esf_p32_eca_fuse"), "p32_eca_fuse",
            "C=%d" % C, nbytes=2 * self._nbytes(x_fast.f32) + self._nbytes(y_slice.f32), launches=2)
        self._post(y_slice.f32, y_slice, label="split")

    def position_attention(self, x_slow, y_slice, alpha, w_down, att, bn):
        """Same composition as Plan.position_attention; the projection GEMM runs on the split operands, the fused
        attention kernel writes FP32 into the concat slice, which is then split into planes."""
        B, T, H, W, C = x_slow.shape
        wd = host64(w_down).reshape(w_down.shape[0], C)
        d = wd.shape[0]
        if d not in (8, 16, 32, 64, 128):
            raise NotImplementedError("head dim %d is not part of the FP32-accurate plan (8, 16, 32, 64, 128)" % d)
        mats, biases = [wd], [torch.zeros(d, dtype=torch.float64)]
        for conv in (att.query_conv, att.key_conv, att.value_conv):
            wc = host64(conv.weight).reshape(conv.weight.shape[0], d)
            bc = host64(conv.bias)
            if wc.shape[0] < d:
                # SpatialAttention(reduction > 1) (wdf_attention_helper.py:17-26): query / key have d / reduction
                # channels.  Zero rows up to d leave every logit q.k unchanged and keep the kernels' one head dim.
                # (No cfg key reaches this: every call site of the reference hard-codes reduction = 1.)
                assert conv is not att.value_conv, "value_conv keeps all channels"
                pad = d - wc.shape[0]
                wc = torch.cat([wc, torch.zeros(pad, d, dtype=wc.dtype)], 0)
                bc = torch.cat([bc, torch.zeros(pad, dtype=bc.dtype)], 0)
            mats.append(wc @ wd)
            biases.append(bc)
        w_all = torch.cat(mats, 0).reshape(4 * d, C, 1, 1, 1)
        b_all = torch.cat(biases, 0)
        proj = super().act(B, T, H, W, 4 * d, dtype=torch.float32)
        self.conv(x_slow, proj, w_all, b_all)
        L = rt.lib()
        N = T * H * W
        scale, shift = bn_affine(bn)
        sc, sh = self.tensor(scale), self.tensor(shift)
        gamma = float(att.gamma.detach().float().item())
        yv = rt.view(y_slice.f32)
        self.keep.append(yv)
        if d <= 64 and self.attn_impl == "tcgen05":
            # tensor cores at FP32 accuracy: the logits are hi/lo split already; P and V become FP16 pairs too
            # (esf_attn_tc_create_split: O += P_hi V_hi + P_lo V_hi + P_hi V_lo, row sums in FP32 registers)
            nbytes = int(L.esf_attn_tc_pack_bytes(B, N, d))
            nlo = int(L.esf_attn_tc_vlo_bytes(B, N, d))
            if nbytes < 0 or nlo < 0:
                rt.check(min(nbytes, nlo), "esf_attn_tc_pack_bytes")
            packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            vlo = torch.empty(nlo, dtype=torch.uint8, device=self.device)
            self.keep += [packed, vlo]
            h = ctypes.c_void_p()
            rt.check(L.esf_attn_tc_create_split(packed.data_ptr(), vlo.data_ptr(), B, T, H, W, d, gamma, sc.data_ptr(),
                                                sh.data_ptr(), alpha, ctypes.byref(yv), ctypes.byref(h)),
                     "esf_attn_tc_create_split")
            self.handles.append(h)

            def pack(s):
                rt.check(L.esf_attn_tc_pack(proj.data_ptr(), B, N, d, rt.F16, packed.data_ptr(), s), "esf_attn_tc_pack")
                rt.check(L.esf_attn_tc_pack_vlo(proj.data_ptr(), B, N, d, vlo.data_ptr(), s), "esf_attn_tc_pack_vlo")

            self._add(pack, "attn_pack", "N=%d d=%d split" % (N, d), nbytes=self._nbytes(proj) + nbytes + nlo, launches=2)
            self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "attention",
                      "N=%d d=%d split" % (N, d), flops=8.0 * B * N * N * d, exps=float(B) * N * N,
                      nbytes=nbytes + nlo + self._nbytes(y_slice.f32))
        else:
            # d = 128 has no hi/lo logit split on the tensor cores (the split Q tile does not fit in shared memory):
            # FP32 flash attention on the CUDA cores (N = 1 568 there)
            self._add(lambda s: rt.check(
                L.esf_p32_attention(proj.data_ptr(), B, T, H, W, d, gamma, sc.data_ptr(), sh.data_ptr(), alpha,
                                    ctypes.byref(yv), s), "esf_p32_attention"), "p32_attention", "N=%d d=%d" % (N, d),
                flops=4.0 * B * N * N * d, exps=float(B) * N * N, nbytes=self._nbytes(proj) + self._nbytes(y_slice.f32))
        self._post(y_slice.f32, y_slice, label="split")

    def head(self, xs, weight, bias, act):
        B = xs[0].shape[0]
        cin = sum(x.shape[4] for x in xs)
        K = weight.shape[0]
        feat = torch.empty((B, cin), dtype=torch.float32, device=self.device)
        out = torch.empty((B, K), dtype=torch.float32, device=self.device)
        w, b = self.tensor(weight), self.tensor(bias)
        self.keep += [feat, out]
        L = rt.lib()
        off = 0
        for x in xs:
            xv = rt.view(x.f32)
            self.keep.append(xv)
            self._add(lambda s, xv=xv, off=off: rt.check(L.esf_p32_head_pool(ctypes.byref(xv), feat.data_ptr(), cin, off, s),
                                                         "esf_p32_head_pool"), "p32_head_pool", "",
                      nbytes=self._nbytes(x.f32))
            off += x.shape[4]
        from .engine import head_fc_launches
        self._add(lambda s: rt.check(
            L.esf_head_fc(feat.data_ptr(), B, cin, cin, w.data_ptr(), b.data_ptr(), K, act, out.data_ptr(), K, s),
            "esf_head_fc"), "head_fc", "", flops=2.0 * B * cin * K, nbytes=self._nbytes(feat, w, out),
            launches=head_fc_launches(B, cin, K, act))
        self.out = out
        return out

    def head_positions(self, xs, pool_sizes, weight, bias, act):
        raise NotImplementedError("fully-convolutional testing is not part of the FP32-accurate plan")

    def nonlocal_block(self, x, y, nln, group=1):
        """Nonlocal.forward (nonlocal_helper.py:105-148) at FP32 accuracy: theta / phi / g / out are split-operand convs;
        the two matrix products are per-clip-weight GEMMs on split operands as well -- theta and the normalised affinity
        are A operands ([hi | lo | hi] planes), the clip's phi rows and transposed g rows are the weight operands
        ([hi | hi | lo]); the affinity is materialised in FP32 and normalised by esf_p32_row_softmax.  (Not fused: an
        accuracy mode.)"""
        if group > 1:
            raise NotImplementedError("NONLOCAL.GROUP > 1 is not part of the FP32-accurate plan")
        L = rt.lib()
        B, T, H, W, C = x.shape
        d = nln.dim_inner
        dp = (d + 7) // 8 * 8

        def wb(conv):
            return host64(conv.weight), host64(conv.bias)

        if nln.use_pool:
            ps = [int(v) for v in nln.pool_size]
            Tp, Hp, Wp = (T - ps[0]) // ps[0] + 1, (H - ps[1]) // ps[1] + 1, (W - ps[2]) // ps[2] + 1
            xp = self.act(B, Tp, Hp, Wp, C)
            self.pool(x, xp, tuple(ps), tuple(ps), (0, 0, 0))
        else:
            Tp, Hp, Wp, xp = T, H, W, x
        Nq, Nk = T * H * W, Tp * Hp * Wp
        softmax = nln.instantiation == "softmax"
        if not softmax and nln.instantiation != "dot_product":
            raise NotImplementedError("Unknown norm type {}".format(nln.instantiation))
        theta = self.act(B, T, H, W, d)
        self.conv(x, theta, *wb(nln.conv_theta))
        # phi rows of a clip ARE the [n][k] weight matrix of theta^T phi (k = the 3 d plane channels): allocate them
        # inside a zero-padded (B, n_pad, 3 dp) matrix and view its first Nk rows as the activation
        kc1, kch1, _, npad1 = rt.igemm_geometry(3 * dp, Nk)
        assert kc1 * kch1 == 3 * dp, "3 x dim_inner must fill whole K chunks"
        phi_rows = torch.zeros((B, npad1, 3 * dp), dtype=torch.float16, device=self.device)
        phi32 = torch.empty((B, Tp, Hp, Wp, d), dtype=torch.float32, device=self.device)
        phi_x3 = phi_rows.as_strided((B, Tp, Hp, Wp, 3 * dp), (npad1 * 3 * dp, Hp * Wp * 3 * dp, Wp * 3 * dp, 3 * dp, 1))
        phi = Act32(phi32, phi_x3, dp, 0, d, weight_order=True)
        g32 = torch.empty((B, Tp, Hp, Wp, d), dtype=torch.float32, device=self.device)
        g_x3 = torch.zeros((B, Tp, Hp, Wp, 3 * dp), dtype=torch.float16, device=self.device)
        g = Act32(g32, g_x3, dp, 0, d, weight_order=True)
        self.keep += [phi_rows, phi32, g32, g_x3]
        self.conv(xp, phi, *wb(nln.conv_phi))
        self.conv(xp, g, *wb(nln.conv_g))
        # affinity S = theta^T phi (FP32), one launch for the whole batch
        S = self.scratch((B, T, H, W, (Nk + 3) // 4 * 4), torch.float32)
        zero1 = self.scratch((npad1,), torch.float32, zero=True)
        self.gemm_rows(theta.x3, S[..., :Nk], phi_rows, zero1, "nl_gemm", "theta.phi x3 N=%dx%d d=%d" % (Nq, Nk, d))
        # normalise into the [hi | lo | hi] planes of the second product's A operand
        nkp = (Nk + 7) // 8 * 8
        P3 = self.scratch((B, T, H, W, 3 * nkp), torch.float16, zero=True)
        scale, mode = (float(d) ** -0.5, 0) if softmax else (1.0 / Nk, 1)
        self._add(lambda s: rt.check(L.esf_p32_row_softmax(S.data_ptr(), B * Nq, Nk, S.shape[4], scale, mode, P3.data_ptr(),
                                                           3 * nkp, nkp, s), "esf_p32_row_softmax"),
                  "p32_nl_softmax", nln.instantiation, nbytes=B * Nq * Nk * 10.0, exps=float(B * Nq * Nk))
        # g^T in weight order: gT3[b][c][plane * nkp + k] = plane(g)[b, k, c], planes (hi, hi, lo)
        kc2, kch2, _, npad2 = rt.igemm_geometry(3 * nkp, d)
        kpad = kc2 * kch2
        gT3 = torch.zeros((B, npad2, kpad), dtype=torch.float16, device=self.device)
        self.keep.append(gT3)
        for pl in range(3):
            src = g_x3.data_ptr() + 2 * pl * dp
            dst = gT3.data_ptr() + 2 * pl * nkp
            self._add(lambda s, src=src, dst=dst: rt.check(
                L.esf_transpose16(src, B, Nk, d, Nk * 3 * dp, 3 * dp, dst, npad2 * kpad, kpad, s), "esf_transpose16"),
                "nl_transpose", "g plane %d" % pl, nbytes=4.0 * B * Nk * d)
        att = self.act(B, T, H, W, d)
        acc = self.scratch((B, T, H, W, d), torch.float32)
        zero2 = self.scratch((npad2,), torch.float32, zero=True)
        self.gemm_rows(P3, acc, gT3, zero2, "nl_gemm", "p.g x3 N=%dx%d d=%d" % (Nq, Nk, d))
        self._post(acc, att, label="split")
        w, bias = fold_conv_bn(nln.conv_out.weight, nln.conv_out.bias, nln.bn)
        self.conv(att, y, w, bias, res=x)
