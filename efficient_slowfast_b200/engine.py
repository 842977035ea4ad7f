"""Launch-plan builder: turns a model's parameters into folded/packed device weights and an ordered list of
libesf_b200 kernel launches over pre-allocated channels-last BF16 activation buffers.

Design (B200-first, not a module-by-module translation of the reference):
  * activations are (B,T,H,W,C) BF16; every torch.cat of the reference is a channel slice of one wider buffer that
    the producers write into directly (concat is free), every eval BatchNorm3d is folded into the preceding conv,
    ReLU / residual adds live in the conv epilogue;
  * a plan is built once per input shape, captured into a CUDA graph and replayed;
  * there is no PyTorch compute on the path -- torch only owns the memory and the stream.
"""
import ctypes
import os

import torch

from . import runtime as rt

BN_EPS_DEFAULT = 1e-5


def host64(t):
    """Parameter / buffer -> FP64 tensor on the HOST.  All weight preparation (BN folding, composition of 1x1x1 convs,
    packing into the kernels' layouts) is plan-build-time work on a few MB and runs on the CPU: one D2H copy per
    parameter and one H2D copy per packed tensor, no element-wise GPU launches -- the first kernels a process launches
    are the forward path's own (the round-1 build issued > 1 000 tiny FP64 ATen kernels per plan)."""
    return t.detach().cpu().to(torch.float64)


def to_device(t, device, dtype):
    """Host tensor -> contiguous device tensor of `dtype`; the conversion happens on the host."""
    return t.detach().to(dtype).contiguous().to(device)


def bn_affine(bn):
    """Eval-mode BatchNorm as y = x * scale + shift (FP64 math on the host)."""
    var = host64(bn.running_var)
    w = host64(bn.weight) if bn.weight is not None else torch.ones_like(var)
    b = host64(bn.bias) if bn.bias is not None else torch.zeros_like(var)
    scale = w / torch.sqrt(var + bn.eps)
    shift = b - host64(bn.running_mean) * scale
    return scale, shift


def fold_conv_bn(weight, conv_bias, bn):
    """(Cout,Cin/g,kT,kH,kW) conv weight (+ optional bias) followed by eval BN -> folded FP64 weight and bias."""
    w = host64(weight)
    cout = w.shape[0]
    b = host64(conv_bias) if conv_bias is not None else torch.zeros(cout, dtype=torch.float64)
    if bn is not None:
        scale, shift = bn_affine(bn)
        w = w * scale.view(-1, 1, 1, 1, 1)
        b = b * scale + shift
    return w, b


def pack_igemm_weight(w, bias, device, dtype=torch.bfloat16):
    """Folded (Cout,Cin,kT,kH,kW) FP64 weight -> 16-bit [n_pad][taps*kchunks*kc] (tap-major, then input channel) and
    FP32 bias [n_pad], the layout esf_conv_igemm_create expects."""
    cout, cin, kt, kh, kw = w.shape
    kc, kchunks, _, n_pad = rt.igemm_geometry(cin, cout)
    cpad = kc * kchunks
    wp = torch.zeros(n_pad, kt * kh * kw, cpad, dtype=torch.float64, device=w.device)
    wp[:cout, :, :cin] = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh * kw, cin)
    bp = torch.zeros(n_pad, dtype=torch.float64, device=w.device)
    bp[:cout] = bias
    return to_device(wp.reshape(n_pad, -1), device, dtype), to_device(bp, device, torch.float32)


STEM_WB = 8  # output columns per banded-GEMM block (csrc/esf_igemm.cu kStemWB)


def pack_stem_band(w, bias, stride_w, device, dtype=torch.bfloat16):
    """Folded (Cout,Cin,kT,kH,kW) stem weight -> band matrix BF16 [n_pad][kT*kH*64] and tiled bias FP32 [n_pad].
    Row n = i*Cout + co (i = output column inside the 8-wide block); column k = (kt*kH + kh)*64 + j where
    (w_in, c) = divmod(j, Cin) is the j-th element of the contiguous input run the block reads and kw = w_in - sW*i."""
    cout, cin, kt, kh, kw = w.shape
    N = STEM_WB * cout
    _, _, _, n_pad = rt.igemm_geometry(64, N)
    band = torch.zeros(n_pad, kt * kh, 64, dtype=torch.float64, device=w.device)
    win = ((STEM_WB - 1) * stride_w + kw) * cin
    assert win <= 64
    wt = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh, kw, cin)          # [co][tap][kw][c]
    for i in range(STEM_WB):
        j0 = stride_w * i * cin
        band[i * cout:(i + 1) * cout, :, j0:j0 + kw * cin] = wt.reshape(cout, kt * kh, kw * cin)
    bt = torch.zeros(n_pad, dtype=torch.float64, device=w.device)
    bt[:N] = bias.repeat(STEM_WB)
    return to_device(band.reshape(n_pad, -1), device, dtype), to_device(bt, device, torch.float32)


def pack_stem_tband(w, bias, WB, stride_w, device, dtype=torch.bfloat16):
    """Folded (Cout,Cin,kT,kH,kW) stem weight -> band matrix [kT*NB][kH*64] and tiled bias FP32 [NB] of the temporal-band
    stem (esf_stem_tband_create), NB = WB*Cout.  Row n = u*NB + i*Cout + co: u is the output frame relative to the oldest
    frame an input frame touches, i.e. the weights of time tap kt = kT-1-u; i = output column inside the WB-wide block.
    Column k = kh*64 + j, (w_in, c) = divmod(j, Cin) the j-th element of the input run the block reads, kw = w_in - sW*i."""
    cout, cin, kt, kh, kw = w.shape
    NB = WB * cout
    assert ((WB - 1) * stride_w + kw) * cin <= 64
    band = torch.zeros(kt, WB, cout, kh, 64, dtype=torch.float64, device=w.device)
    wt = w.to(torch.float64).permute(2, 0, 3, 4, 1).reshape(kt, cout, kh, kw * cin)      # [kt][co][kh][(kw, c)]
    for i in range(WB):
        j0 = stride_w * i * cin
        band[:, i, :, :, j0:j0 + kw * cin] = wt.flip(0)                                   # u = kT-1-kt
    bt = bias.to(torch.float64).repeat(WB)
    return to_device(band.reshape(kt * NB, kh * 64), device, dtype), to_device(bt, device, torch.float32)


def pack_wfold_band(w, bias, WB, stride_w, device, dtype=torch.bfloat16):
    """Folded (Cout,Cin,kT,kH,kW) weight of a thin layer -> band matrix [n_pad][kT*kH*kchunks*64] and tiled bias [n_pad]
    for esf_conv_wfold_create.  Row n = i*Cout + co (i = output column inside the WB-wide block); column
    k = (kt*kH + kh)*kchunks*64 + j where (w_in, c) = divmod(j, Cin) indexes the contiguous input run the block reads
    (starting at input column WB*sW*block - pW) and kw = w_in - sW*i."""
    cout, cin, kt, kh, kw = w.shape
    N = WB * cout
    _, _, _, n_pad = rt.igemm_geometry(64, N)
    win = ((WB - 1) * stride_w + kw) * cin
    kpad = -(-win // 64) * 64
    band = torch.zeros(n_pad, kt * kh, kpad, dtype=torch.float64, device=w.device)
    wt = w.permute(0, 2, 3, 4, 1).reshape(cout, kt * kh, kw * cin).to(torch.float64)   # [co][tap][(kw, c)]
    for i in range(WB):
        j0 = stride_w * i * cin
        band[i * cout:(i + 1) * cout, :, j0:j0 + kw * cin] = wt
    bt = torch.zeros(n_pad, dtype=torch.float64, device=w.device)
    bt[:N] = bias.repeat(WB)
    return to_device(band.reshape(n_pad, -1), device, dtype), to_device(bt, device, torch.float32)


def wfold_block(x, y, res, w_shape, stride, padding, dilation):
    """Width of the output-column block to fold into the GEMM's N for a thin layer, or 0 when the regular implicit
    GEMM should run.  Mirrors the argument checks of esf_conv_wfold_create."""
    cout, cin, kt, kh, kw = w_shape
    if cin > 32 or cin % 8 or stride[0] != 1 or tuple(dilation) != (1, 1, 1):
        return 0
    if x.stride(3) != cin or x.stride(4) != 1:
        return 0
    Wo = y.shape[3]
    sliced = y.stride(3) != cout or (res is not None and res.stride(3) != cout)
    best, best_key = 0, None
    for wb in range(1, 33):
        N = wb * cout
        win = ((wb - 1) * stride[2] + kw) * cin
        if Wo % wb or N > 256 or win > 128:
            continue
        if sliced and (N & (N - 1) or (64 % cout if cout < 64 else cout % 64)):
            continue
        key = (N, -((win + 63) // 64))      # widest MMA first, then fewest K chunks
        if best_key is None or key > best_key:
            best, best_key = wb, key
    return best if best >= 2 else 0     # a block of one column is the plain implicit GEMM


def head_fc_launches(B, cin, K, act):
    """Kernel launches of one esf_head_fc call: the tiled kernel (B >= 8, or rows too long for one block) applies the
    class softmax in a second, tiny kernel."""
    tiled = B >= 8 or (cin + K) * 4 > 48 * 1024
    return 2 if (tiled and act == rt.HEAD_SOFTMAX) else 1


class Plan:
    """Ordered kernel launches + every tensor they touch.  `eager` ops read the caller's input tensors and are
    launched on every forward; `graph` ops only touch plan-owned memory and are replayed from one CUDA graph."""

    def __init__(self, device, precision="bf16"):
        self.device = device
        self.precision = precision
        self.adt = rt.TORCH_DTYPE[precision]      # 16-bit storage format of activations / tensor-core operands
        self.a16 = rt.dtype_code(self.adt)
        self.keep = []        # tensors / ctypes structs that must outlive the plan
        self.ops = []         # callables taking a stream pointer
        self.eager = []       # per op: launched outside the CUDA graph (reads a caller-provided input tensor)
        self.input_override = {}   # plan-owned input data_ptr -> data_ptr of the caller's tensor for this run
        self.stem_routes = {}      # plan-owned input data_ptr -> (packed stem rows, pitch, lpad) of a banded stem
        self.gather = {}           # plan-owned input data_ptr -> (source clip ptr, its frame count, device int32 frame
                                   # index ptr): this run packs the pathway out of ANOTHER clip's frames (slow from fast)
        self.padded = {}           # data_ptr of a channel-padded activation -> (logical C, padded C)
        self.handles = []     # esf_op* to destroy
        self.graph = None
        self.out = None
        self.launches_per_run = 0
        self.buffers = {}     # name -> activation tensor (for tests / debugging)
        self.meta = []        # one dict per op: kind, label, algorithmic flops / bytes / exps, launches
        self.wfold = True            # thin layers (C_in <= 32) as W-folded banded GEMMs
        self.attn_impl = "tcgen05"   # "tcgen05" (TMEM, two-pass) or "mma_sync" (register-resident, online softmax)

    # ---------------------------------------------------------------- memory
    def act(self, B, T, H, W, C, name=None, dtype=None, role=None):
        """role: a hint for the FP32-accurate plan (engine_fp32.PrecisePlan.act); 16-bit activations have one form."""
        dtype = dtype or self.adt
        # 16-bit activations with C >= 8 get a channel pitch that is a multiple of 8 elements (16 bytes), so that every
        # buffer -- also with C = 27, 36, 180, 540 ... of the efficient backbones -- is addressable by TMA and by
        # 16-byte vector accesses; the logical tensor is the [..., :C] view
        Cp = C if (dtype != self.adt or C < 8) else (C + 7) // 8 * 8
        t = torch.empty((B, T, H, W, Cp), dtype=dtype, device=self.device)
        self.keep.append(t)
        if Cp != C:
            self.padded[t.data_ptr()] = (C, Cp)     # channels [C, Cp) of every row are padding owned by this tensor
            t = t[..., :C]
        if name:
            self.buffers[name] = t
        return t

    def arena_bytes(self):
        """Device bytes this plan keeps alive (activations, packed weights, scratch, static inputs)."""
        seen, total = set(), 0
        for t in list(self.keep) + list(getattr(self, "inputs", [])):
            if isinstance(t, torch.Tensor) and t.device.type == "cuda":
                st = t.untyped_storage()
                if st.data_ptr() not in seen:
                    seen.add(st.data_ptr())
                    total += st.nbytes()
        return total

    def tensor(self, t, dtype=torch.float32):
        """Plan-lifetime device copy of a (small) parameter-derived tensor; host tensors are converted on the host."""
        if t.device.type == "cpu":
            t = to_device(t, self.device, dtype)
        else:
            t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.keep.append(t)
        return t

    def _add(self, fn, kind, label="", flops=0, nbytes=0, exps=0, launches=1, eager=False):
        """`eager` ops read the caller's input tensors (their source pointer can change per call): they stay outside
        the CUDA graph and are launched on every run, ahead of the graph replay."""
        self.ops.append(fn)
        self.eager.append(bool(eager))
        self.meta.append(dict(kind=kind, label=label, flops=float(flops), bytes=float(nbytes), exps=float(exps),
                              launches=launches))
        self.launches_per_run += launches

    @staticmethod
    def _nbytes(*tensors):
        return sum(t.numel() * t.element_size() for t in tensors if t is not None)

    # ---------------------------------------------------------------- ops
    def conv_igemm(self, x, y, w_folded, bias, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1), act=rt.ACT_NONE,
                   res=None, out_dtype=None):
        out_dtype = rt.dtype_code(y)
        wp, bp = pack_igemm_weight(w_folded, bias, self.device, self.adt)
        self.keep += [wp, bp]
        kt, kh, kw = w_folded.shape[2:]
        d = rt.EsfConvDesc(rt.view(x), rt.view(y), rt.view(res) if res is not None else rt.null_view(),
                           wp.data_ptr(), bp.data_ptr(), kt, kh, kw, *stride, *padding, *dilation, 1, act, out_dtype)
        h = ctypes.c_void_p()
        rt.check(rt.lib().esf_conv_igemm_create(ctypes.byref(d), ctypes.byref(h)), "esf_conv_igemm_create")
        self.handles.append(h)
        L = rt.lib()
        cout, cin = w_folded.shape[:2]
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "conv_igemm",
                  "%dx%dx%d s%s %d->%d @%s" % (kt, kh, kw, "".join(map(str, stride)), cin, cout, tuple(y.shape[1:4])),
                  flops=2.0 * m * cout * cin * kt * kh * kw, nbytes=self._nbytes(x, y, res) + wp.numel() * 2)

    def gemm_rows(self, x, y, w_rows, bias, kind, label):
        """y[b, pos, n] = sum_c x[b, pos, c] * w_rows[b, n, c] + bias[n] on the implicit-GEMM kernel with one
        DEVICE-RESIDENT weight matrix per clip, in the kernel's own layout ([B][n_pad][kchunks * kc] 16-bit rows, FP32
        bias [n_pad]) -- used where the "weights" are activations of the same clip (Non-local block)."""
        cin, cout = x.shape[4], y.shape[4]
        kc, kchunks, _, n_pad = rt.igemm_geometry(cin, cout)
        assert tuple(w_rows.shape) == (x.shape[0], n_pad, kc * kchunks) and w_rows.is_contiguous()
        assert w_rows.dtype == self.adt and bias.numel() == n_pad and bias.dtype == torch.float32
        d = rt.EsfConvDesc(rt.view(x), rt.view(y), rt.null_view(), w_rows.data_ptr(), bias.data_ptr(), 1, 1, 1,
                           1, 1, 1, 0, 0, 0, 1, 1, 1, 1, rt.ACT_NONE, rt.dtype_code(y))
        h = ctypes.c_void_p()
        L = rt.lib()
        rt.check(L.esf_gemm_clip_weights_create(ctypes.byref(d), ctypes.byref(h)), "esf_gemm_clip_weights_create")
        self.handles.append(h)
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), kind, label,
                  flops=2.0 * m * cout * cin, nbytes=self._nbytes(x, y) + x.shape[0] * cout * cin * 2)

    def scratch(self, shape, dtype, zero=False):
        """Plan-lifetime scratch tensor shared by every op that asks for the same (shape, dtype): the launches of a
        plan are serialised on one stream, so the Non-local blocks of a model reuse one affinity matrix."""
        key = (tuple(shape), dtype, bool(zero))
        if not hasattr(self, "_scratch"):
            self._scratch = {}
        if key not in self._scratch:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self.keep.append(t)
            self._scratch[key] = t
        return self._scratch[key]

    def nonlocal_block(self, x, y, nln, group=1):
        """Nonlocal.forward (nonlocal_helper.py:105-148) + the temporal group folding of ResStage.forward
        (resnet_helper.py:541-560): y = x + bn(conv_out(normalise(theta^T phi) g^T)).

        theta / phi / g / out are ordinary implicit GEMMs (bias in the epilogue; BN folded into conv_out, the residual
        add in its epilogue).  The two matrix products are implicit GEMMs as well (esf_gemm_clip_weights_create: one
        launch for the whole batch, the weight tile of an M tile is selected by its clip), whose weight operand is
        the clip's own phi rows ([N_keys][d], exactly the kernel's [n][k] layout) and g^T rows
        ([d][N_keys], written by esf_transpose16); the affinity matrix is materialised once in FP32, normalised by
        esf_row_softmax (softmax with d^-0.5, or the 1/N_keys of "dot_product") into the 16-bit A operand of the
        second product."""
        L = rt.lib()
        if group > 1:   # (b, t) -> (b * group, t / group): a pure view in channels-last memory
            def fold(t):
                B, T, H, W, C = t.shape
                assert T % group == 0 and t.stride(0) == T * t.stride(1)
                return t.as_strided((B * group, T // group, H, W, C),
                                    (t.stride(1) * (T // group), t.stride(1), t.stride(2), t.stride(3), 1),
                                    t.storage_offset())
            x, y = fold(x), fold(y)
        B, T, H, W, C = x.shape
        d = nln.dim_inner
        f64 = torch.float64

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
        # Precision: with near-uniform attention the block output is a large per-channel constant plus a small
        # variation, and the BN behind conv_out removes the constant -- 16-bit rounding of `att` would then be
        # amplified by |mean| / std.  conv_out is linear, so a per-channel offset mu can be subtracted in the FP32
        # epilogue of the last product (its bias) and added back through conv_out's bias: W (att - mu) + (W mu + b).
        # mu = least-squares solution of W mu = running_mean - b, the att-space mean the BN statistics imply.
        w_out = host64(nln.conv_out.weight).reshape(C, d)
        rhs = (host64(nln.bn.running_mean) - host64(nln.conv_out.bias)).reshape(C, 1)
        mu = torch.linalg.lstsq(w_out, rhs).solution.reshape(d)
        _, _, _, npad_d = rt.igemm_geometry(Nk if softmax else d, d)
        bias2_h = torch.zeros(npad_d, dtype=torch.float32)
        bias2_h[:d] = (-mu).to(torch.float32)
        mu = -bias2_h[:d].to(f64)                    # the value actually subtracted (FP32-rounded)
        bias2 = self.tensor(bias2_h)
        att = self.act(B, T, H, W, d)

        def transposed(t, name):
            """(B, Tp, Hp, Wp, d) rows -> [B][n_pad][k_pad] = t^T per clip, zero padded: the [n][k] weight layout of a
            GEMM that contracts over the keys."""
            kc, kch, _, npad = rt.igemm_geometry(Nk, d)
            kpad = kc * kch
            tT = torch.zeros((B, npad, kpad), dtype=self.adt, device=self.device)
            self.keep.append(tT)
            self._add(lambda s: rt.check(L.esf_transpose16(t.data_ptr(), B, Nk, d, t.stride(0), t.stride(3), tT.data_ptr(),
                                                           npad * kpad, kpad, s), "esf_transpose16"),
                      "nl_transpose", name, nbytes=2 * self._nbytes(t))
            return tT

        if softmax:
            # phi rows of a clip = the [n][k] weight matrix of theta^T phi: rows padded to the GEMM's n_pad
            kc1, kch1, _, npad1 = rt.igemm_geometry(d, Nk)
            assert kc1 * kch1 == d, "dim_inner must fill whole K chunks"
            theta = self.act(B, T, H, W, d)
            self.conv(x, theta, *wb(nln.conv_theta))
            phi_rows = torch.zeros((B, npad1, d), dtype=self.adt, device=self.device)
            phi = phi_rows.as_strided((B, Tp, Hp, Wp, d), (npad1 * d, Hp * Wp * d, Wp * d, d, 1))
            g = self.act(B, Tp, Hp, Wp, d)
            self.keep += [phi_rows]
            self.conv(xp, phi, *wb(nln.conv_phi))
            self.conv(xp, g, *wb(nln.conv_g))
            gT = transposed(g, "g")
            Sbuf = self.scratch((B, T, H, W, (Nk + 3) // 4 * 4), torch.float32)
            S = Sbuf[..., :Nk]
            Pbuf = self.scratch((B, T, H, W, (Nk + 7) // 8 * 8), self.adt)
            P = Pbuf[..., :Nk]
            zero1 = self.scratch((npad1,), torch.float32, zero=True)
            self.gemm_rows(theta, S, phi_rows, zero1, "nl_gemm", "theta.phi N=%dx%d d=%d" % (Nq, Nk, d))
            self._add(lambda s: rt.check(L.esf_row_softmax(Sbuf.data_ptr(), B * Nq, Nk, Sbuf.shape[4], float(d) ** -0.5, 0,
                                                           self.a16, Pbuf.data_ptr(), Pbuf.shape[4], s), "esf_row_softmax"),
                      "nl_softmax", nln.instantiation, nbytes=self._nbytes(S, P), exps=float(B * Nq * Nk))
            self.gemm_rows(P, att, gT, bias2, "nl_gemm", "p.g N=%dx%d d=%d" % (Nq, Nk, d))
        else:
            # "dot_product" has no non-linearity between the two products:  (theta^T phi / Nk) g^T = theta^T (phi g^T / Nk),
            # a d x d matrix per clip instead of the Nq x Nk affinity -- 2 Nk d^2 + 2 Nq d^2 FLOP instead of 4 Nq Nk d,
            # and nothing of size Nq x Nk is ever written.  M^T[c][c'] = sum_p (g[p][c] / Nk) phi[p][c'] is itself a
            # per-clip-weights GEMM over the transposed rows, and its 16-bit output IS the [n][k] weight matrix of
            # att = theta M.  (The 1 / Nk goes into the g conv so that M stays O(1) in 16 bits.)
            assert rt.igemm_geometry(d, d)[3] == d, "dim_inner must be a whole N tile"
            theta = self.act(B, T, H, W, d)
            self.conv(x, theta, *wb(nln.conv_theta))
            phi = self.act(B, Tp, Hp, Wp, d)
            g = self.act(B, Tp, Hp, Wp, d)
            self.conv(xp, phi, *wb(nln.conv_phi))
            wg, bg = wb(nln.conv_g)
            self.conv(xp, g, wg / Nk, bg / Nk)
            phiT, gT = transposed(phi, "phi"), transposed(g, "g")
            MT = torch.zeros((B, 1, 1, d, d), dtype=self.adt, device=self.device)
            self.keep.append(MT)
            zero_d = self.scratch((d,), torch.float32, zero=True)
            kpad = gT.shape[2]
            gT_act = gT.as_strided((B, 1, 1, d, Nk), (gT.stride(0), d * kpad, d * kpad, kpad, 1))
            self.gemm_rows(gT_act, MT, phiT, zero_d, "nl_gemm", "g^T.phi d=%d Nk=%d" % (d, Nk))
            self.gemm_rows(theta, att, MT.reshape(B, d, d), bias2, "nl_gemm", "theta.M N=%d d=%d" % (Nq, d))
        w, bias = fold_conv_bn(nln.conv_out.weight, nln.conv_out.bias, nln.bn)
        bias = bias + w.reshape(C, d) @ mu
        self.conv(att, y, w, bias, res=x)

    def conv_wfold(self, x, y, w_folded, bias, WB, stride=(1, 1, 1), padding=(0, 0, 0), act=rt.ACT_NONE, res=None):
        """Thin-layer (C_in <= 32) conv as a W-folded banded GEMM: see esf_conv_wfold_create."""
        wp, bp = pack_wfold_band(w_folded, bias, WB, stride[2], self.device, self.adt)
        self.keep += [wp, bp]
        kt, kh, kw = w_folded.shape[2:]
        d = rt.EsfConvDesc(rt.view(x), rt.view(y), rt.view(res) if res is not None else rt.null_view(),
                           wp.data_ptr(), bp.data_ptr(), kt, kh, kw, *stride, *padding, 1, 1, 1, 1, act,
                           rt.dtype_code(y))
        h = ctypes.c_void_p()
        L = rt.lib()
        rt.check(L.esf_conv_wfold_create(ctypes.byref(d), WB, ctypes.byref(h)), "esf_conv_wfold_create")
        self.handles.append(h)
        cout, cin = w_folded.shape[:2]
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "conv_wfold",
                  "%dx%dx%d s%s %d->%d @%s wb%d" % (kt, kh, kw, "".join(map(str, stride)), cin, cout,
                                                   tuple(y.shape[1:4]), WB),
                  flops=2.0 * m * cout * cin * kt * kh * kw, nbytes=self._nbytes(x, y, res) + wp.numel() * 2)

    @staticmethod
    def _aligned(t):
        """True when a view can be addressed by TMA / 16-byte vector accesses."""
        if t is None:
            return True
        q = 16 // t.element_size()
        return t.data_ptr() % 16 == 0 and all(st % q == 0 for st in t.stride()[:4])

    def conv(self, x, y, w_folded, bias, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1), groups=1,
             act=rt.ACT_NONE, res=None, out_dtype=None):
        """Conv3d + folded BN (+ residual) + activation: tensor-core implicit GEMM when the layer is dense and its
        views are 16-byte addressable, CUDA-core direct conv otherwise (grouped / depthwise / odd channel counts)."""
        if groups == 1 and x.shape[4] >= 8 and self._aligned(x) and self._aligned(y) and self._aligned(res):
            wb = wfold_block(x, y, res, w_folded.shape, stride, padding, dilation) if self.wfold else 0
            if wb:
                return self.conv_wfold(x, y, w_folded, bias, wb, stride, padding, act, res)
            return self.conv_igemm(x, y, w_folded, bias, stride, padding, dilation, act, res, out_dtype)
        cout, cin_g = w_folded.shape[:2]
        if (not self._aligned(y) and self._aligned(x) and self._aligned(res) and x.shape[4] >= 8 and cout >= 8
                and groups <= 8 and groups != x.shape[4] and y.dtype == self.adt and (out_dtype in (None, rt.dtype_code(y)))
                and y.data_ptr() % 4 == 0 and all(st % 2 == 0 for st in y.stride()[:4])):
            # The only obstacle to the tensor-core path is an output slice at a channel offset that is not a multiple of
            # 8 (ShuffleNet: the fast pathway's 15 -> 60 grouped conv lands behind the 60 fused channels of the concat
            # buffer; 0.72 ms on the generic CUDA-core kernel against 0.03 ms of HBM time): run the GEMM into an aligned
            # buffer of its own and copy the rows into the slice.
            tmp = self.act(*y.shape[:4], cout)
            # (with these preconditions the recursion always ends on the implicit GEMM: dense, per-group or block-diagonal)
            self.conv(x, tmp, w_folded, bias, stride, padding, dilation, groups, act, res, out_dtype)
            return self.shuffle_concat(tmp, None, 1, y)
        if (groups > 1 and cin_g % 8 == 0 and (cout // groups) % 8 == 0 and cin_g >= 16 and self._aligned(x)
                and self._aligned(y) and self._aligned(res)):
            # grouped dense conv (ShuffleNet's grouped 1x1x1): one implicit GEMM per group on channel slices
            cout_g = cout // groups
            for g in range(groups):
                self.conv_igemm(x[..., g * cin_g:(g + 1) * cin_g], y[..., g * cout_g:(g + 1) * cout_g],
                                w_folded[g * cout_g:(g + 1) * cout_g], bias[g * cout_g:(g + 1) * cout_g], stride,
                                padding, dilation, act, None if res is None else res[..., g * cout_g:(g + 1) * cout_g],
                                out_dtype)
            return
        if (1 < groups <= 8 and cin_g * groups >= 8 and x.shape[4] != groups and self._aligned(x) and self._aligned(y)
                and self._aligned(res)):
            # grouped conv whose group slices are not 16-byte addressable (ShuffleNet g3: 180 / 18 / 5 channels per
            # group): one dense GEMM over all input channels with a block-diagonal weight.  The layer is HBM bound, the
            # `groups`-fold MACs on structural zeros cost nothing next to the CUDA-core fallback.
            cout_g = cout // groups
            wd = torch.zeros((cout, cin_g * groups) + tuple(w_folded.shape[2:]), dtype=w_folded.dtype,
                             device=w_folded.device)
            for g in range(groups):
                wd[g * cout_g:(g + 1) * cout_g, g * cin_g:(g + 1) * cin_g] = w_folded[g * cout_g:(g + 1) * cout_g]
            return self.conv_igemm(x, y, wd, bias, stride, padding, dilation, act, res, out_dtype)
        self.conv_direct(x, y, w_folded, bias, stride, padding, dilation, groups, act, res, out_dtype)
        why = [n for n, t in (("x", x), ("y", y), ("res", res)) if not self._aligned(t)]
        self.meta[-1]["why_direct"] = ("groups=%d " % groups if groups != 1 else "") + \
            ("cin<8 " if x.shape[4] < 8 else "") + ("unaligned " + ",".join(
                "%s(off %d, strides %s)" % (n, (t.data_ptr() % 16) // 2, tuple(t.stride()[:4]))
                for n, t in (("x", x), ("y", y), ("res", res)) if n in why) if why else "")

    def conv_direct(self, x, y, w_folded, bias, stride=(1, 1, 1), padding=(0, 0, 0), dilation=(1, 1, 1), groups=1,
                    act=rt.ACT_NONE, res=None, out_dtype=None):
        out_dtype = rt.dtype_code(y)
        wd = self.tensor(w_folded.permute(0, 2, 3, 4, 1))  # [Cout][kT][kH][kW][Cin/g]
        bd = self.tensor(bias)
        kt, kh, kw = w_folded.shape[2:]
        d = rt.EsfConvDesc(rt.view(x), rt.view(y), rt.view(res) if res is not None else rt.null_view(),
                           wd.data_ptr(), bd.data_ptr(), kt, kh, kw, *stride, *padding, *dilation, groups, act,
                           out_dtype)
        self.keep.append(d)
        L = rt.lib()
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        depthwise = groups == x.shape[4] == y.shape[4] and kw in (3, 5)
        C = x.shape[4]
        pad_of = lambda t: self.padded.get(t.data_ptr()) if t is not None else (C, (C + 7) // 8 * 8)
        if (depthwise and C % 8 != 0 and C >= 8 and tuple(dilation) == (1, 1, 1) and stride[2] in (1, 2)
                and all(pad_of(t) == (C, (C + 7) // 8 * 8) for t in (x, y, res)) and y.dtype == x.dtype == self.adt):
            # rows padded to a multiple of 8 channels by act(): 16-byte vectors over the padded width (zero weights)
            cp = (C + 7) // 8 * 8
            self._add(lambda s, d=d: rt.check(L.esf_dwconv_padded(ctypes.byref(d), cp, s), "esf_dwconv_padded"),
                      "dwconv", "%dx%dx%d s%d%d%d g%d %d->%d pad%d @%s" % (kt, kh, kw, *stride, groups, C, C, cp,
                                                                              tuple(x.shape[1:4])),
                      flops=2.0 * m * C * kt * kh * kw, nbytes=self._nbytes(x, y, res) + wd.numel() * 4)
            return
        cout = y.shape[4]
        call = lambda s, d=d: rt.check(L.esf_conv_direct(ctypes.byref(d), s), "esf_conv_direct")
        if (groups == 1 and (kt, kh, kw) == (1, 1, 1) and cout >= 8 and cout % 8 != 0 and y.dtype == self.adt
                and self.padded.get(y.data_ptr()) == (cout, (cout + 7) // 8 * 8)):
            # y is a whole padded allocation (not a concat slice): the kernel may write the row padding too
            cp = (cout + 7) // 8 * 8
            call = lambda s, d=d, cp=cp: rt.check(L.esf_pointwise_padded(ctypes.byref(d), cp, s), "esf_pointwise_padded")
        self._add(call,
                  "dwconv" if depthwise else "conv_direct",
                  "%dx%dx%d s%d%d%d g%d %d->%d @%s" % (kt, kh, kw, *stride, groups, x.shape[4], y.shape[4],
                                                       tuple(x.shape[1:4])),
                  flops=2.0 * m * y.shape[4] * (x.shape[4] // groups) * kt * kh * kw,
                  nbytes=self._nbytes(x, y, res) + wd.numel() * 4)

    def stem_conv(self, x_nc, y, w_folded, bias, stride, padding, act=rt.ACT_RELU):
        """x_nc: FP32 (B,Cin,T,H,W) contiguous plan-owned input buffer."""
        ws = self.tensor(w_folded.permute(2, 3, 4, 1, 0))  # [kT][kH][kW][Cin][Cout]
        bs = self.tensor(bias)
        B, Cin, T, H, W = x_nc.shape
        cout = w_folded.shape[0]
        kt, kh, kw = w_folded.shape[2:]
        yv = rt.view(y)
        self.keep.append(yv)
        L = rt.lib()
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        self._add(lambda s: rt.check(
            L.esf_stem_conv(self._in_ptr(x_nc), B, Cin, T, H, W, ws.data_ptr(), bs.data_ptr(), cout, kt, kh, kw, *stride,
                            *padding, act, ctypes.byref(yv), s), "esf_stem_conv"), "stem_conv",
            "%dx%dx%d %d->%d" % (kt, kh, kw, Cin, cout), flops=2.0 * m * cout * Cin * kt * kh * kw,
            nbytes=self._nbytes(x_nc, y), eager=True)

    def _in_ptr(self, x_nc):
        p = x_nc.data_ptr()
        return self.input_override.get(p, p)

    def stem(self, x_nc, y, w_folded, bias, stride, padding, act=rt.ACT_RELU):
        """Stem conv + folded BN + ReLU.  Tensor-core banded GEMM when the geometry allows (temporal stride 1,
        window <= 64 elements, dense output), else the generic CUDA-core stem kernel."""
        B, Cin, T, H, W = x_nc.shape
        cout = w_folded.shape[0]
        kt, kh, kw = w_folded.shape[2:]
        geo = rt.stem_geometry(W, Cin, kw, stride[2], padding[2]) if stride[0] == 1 else None
        if geo is None or not y.is_contiguous():
            return self.stem_conv(x_nc, y, w_folded, bias, stride, padding, act)
        pitch, lpad, _ = geo
        L = rt.lib()
        xp = torch.empty((B, T, H, pitch), dtype=self.adt, device=self.device)
        yv = rt.view(y)
        h = ctypes.c_void_p()
        # kT > 1: the time taps are folded into N and the input frames stream past a resident M tile (csrc/esf_igemm.cu,
        # "temporal-band stem"); ESF_STEM_TBAND=0 keeps the banded stem for A/B runs
        twb = rt.stem_tband_wb(W, Cin, cout, kt, kh, kw, stride[2], padding[2]) \
            if os.environ.get("ESF_STEM_TBAND", "1") != "0" else 0
        if twb:
            wb, bt = pack_stem_tband(w_folded, bias, twb, stride[2], self.device, self.adt)
            rt.check(L.esf_stem_tband_create(xp.data_ptr(), B, Cin, T, H, W, pitch, wb.data_ptr(), bt.data_ptr(), cout,
                                             kt, kh, kw, stride[1], stride[2], padding[0], padding[1], padding[2], act,
                                             ctypes.byref(yv), ctypes.byref(h)), "esf_stem_tband_create")
        else:
            wb, bt = pack_stem_band(w_folded, bias, stride[2], self.device, self.adt)
            rt.check(L.esf_stem_igemm_create(xp.data_ptr(), B, Cin, T, H, W, pitch, wb.data_ptr(), bt.data_ptr(), cout,
                                             kt, kh, kw, stride[1], stride[2], padding[0], padding[1], padding[2], act,
                                             ctypes.byref(yv), ctypes.byref(h)), "esf_stem_igemm_create")
        self.keep += [xp, wb, bt]
        self.handles.append(h)
        m = y.shape[0] * y.shape[1] * y.shape[2] * y.shape[3]
        self.stem_routes[x_nc.data_ptr()] = (xp, pitch, lpad)   # the uint8 frame route writes xp itself (frames.py)
        def pack(s):
            g = self.gather.get(x_nc.data_ptr())
            if g is not None:       # frames of this pathway are read out of another pathway's clip (slow from fast)
                src, tsrc, idx = g
                rt.check(L.esf_stem_pack_gather(src, B, Cin, tsrc, H, W, idx, T, pitch, lpad, self.a16, xp.data_ptr(), s),
                         "esf_stem_pack_gather")
            else:
                rt.check(L.esf_stem_pack(self._in_ptr(x_nc), B, Cin, T, H, W, pitch, lpad, self.a16, xp.data_ptr(), s),
                         "esf_stem_pack")

        self._add(pack, "stem_pack", "", nbytes=self._nbytes(x_nc, xp), eager=True)
        self.meta[-1]["is_pack"] = True
        self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "stem_igemm",
                  "%dx%dx%d %d->%d %s" % (kt, kh, kw, Cin, cout, "t-band" if twb else "banded"), flops=2.0 * m * cout * Cin * kt * kh * kw,
                  nbytes=self._nbytes(xp, y) + wb.numel() * 2)

    def pool(self, x, y, kernel, stride, padding, is_avg=False, act=rt.ACT_NONE):
        xv, yv = rt.view(x), rt.view(y)
        self.keep += [xv, yv]
        L = rt.lib()
        self._add(lambda s: rt.check(
            L.esf_pool3d(ctypes.byref(xv), ctypes.byref(yv), *kernel, *stride, *padding, int(is_avg), act, s),
            "esf_pool3d"), "pool", "%s C=%d" % (tuple(kernel), x.shape[4]), nbytes=self._nbytes(x, y))

    def shuffle_concat(self, a, b, groups, y):
        av, yv = rt.view(a), rt.view(y)
        bv = rt.view(b) if b is not None else rt.null_view()
        self.keep += [av, bv, yv]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_shuffle_concat(ctypes.byref(av), ctypes.byref(bv), groups, ctypes.byref(yv), s),
                                     "esf_shuffle_concat"), "shuffle", "g%d" % groups, nbytes=2 * self._nbytes(y))

    def eltwise_add(self, a, b, y, act=rt.ACT_NONE):
        av, bv, yv = rt.view(a), rt.view(b), rt.view(y)
        self.keep += [av, bv, yv]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_eltwise_add(ctypes.byref(av), ctypes.byref(bv), ctypes.byref(yv), act, s),
                                     "esf_eltwise_add"), "add", "", nbytes=3 * self._nbytes(y))

    def squeeze_excite(self, x, y, se):
        """SqueezeExcite (ghostnet_helper.py:34-52): global mean -> 1x1x1(+bias) -> ReLU -> 1x1x1(+bias) -> hard-sigmoid
        -> channel scale.  Reuses the head kernels for the pooled MLP."""
        B, C = x.shape[0], x.shape[4]
        R = se.conv_reduce.out_channels
        feat = torch.empty((B, C), dtype=torch.float32, device=self.device)
        hid = torch.empty((B, R), dtype=torch.float32, device=self.device)
        gate = torch.empty((B, C), dtype=torch.float32, device=self.device)
        w1, b1 = self.tensor(se.conv_reduce.weight.reshape(R, C)), self.tensor(se.conv_reduce.bias)
        w2, b2 = self.tensor(se.conv_expand.weight.reshape(C, R)), self.tensor(se.conv_expand.bias)
        xv, yv, nv = rt.view(x), rt.view(y), rt.null_view()
        self.keep += [feat, hid, gate, xv, yv, nv]
        L = rt.lib()
        self.global_mean(x, feat, C, 0, "se_pool")
        self._add(lambda s: rt.check(L.esf_head_fc(feat.data_ptr(), B, C, C, w1.data_ptr(), b1.data_ptr(), R, rt.HEAD_RELU,
                                                   hid.data_ptr(), R, s), "esf_head_fc"), "se_fc", "")
        self._add(lambda s: rt.check(L.esf_head_fc(hid.data_ptr(), B, R, R, w2.data_ptr(), b2.data_ptr(), C,
                                                   rt.HEAD_HARD_SIGMOID, gate.data_ptr(), C, s), "esf_head_fc"), "se_fc", "")
        self._add(lambda s: rt.check(L.esf_channel_scale(ctypes.byref(xv), gate.data_ptr(), ctypes.byref(yv), s),
                                     "esf_channel_scale"), "se_scale", "", nbytes=2 * self._nbytes(x))

    def global_mean(self, x, feat, feat_stride, feat_off, kind="head_pool"):
        """feat[:, feat_off : feat_off + C] = mean over (T, H, W) of x: the two-pass esf_global_mean for large maps, the
        single-block head kernel for the small ones."""
        L = rt.lib()
        xv = rt.view(x)
        self.keep.append(xv)
        B, T, H, W, C = x.shape
        if T * H * W >= 1024 and (C % 8 == 0 or C <= 256):
            scratch = self.scratch((int(L.esf_global_mean_scratch_floats(B, C)),), torch.float32)
            self._add(lambda s: rt.check(L.esf_global_mean(ctypes.byref(xv), scratch.data_ptr(), feat.data_ptr(),
                                                           feat_stride, feat_off, s), "esf_global_mean"),
                      kind, "two-pass", nbytes=self._nbytes(x), launches=2)
        else:
            assert feat_off == 0
            nv = rt.null_view()
            self.keep.append(nv)
            self._add(lambda s: rt.check(L.esf_head_pool(ctypes.byref(xv), ctypes.byref(nv), feat.data_ptr(), s),
                                         "esf_head_pool"), kind, "", nbytes=self._nbytes(x))

    def eca_fuse(self, x_fast, y_slice, alpha, eca_weight, bn):
        """MaxPool(alpha,1,1) -> ECA -> BN -> ReLU -> concat slice (custom_video_model_builder.py:131-135)."""
        scale, shift = bn_affine(bn)
        w = self.tensor(eca_weight.reshape(-1))
        sc, sh = self.tensor(scale), self.tensor(shift)
        B, C = x_fast.shape[0], x_fast.shape[4]
        L = rt.lib()
        scratch = torch.empty(int(L.esf_eca_scratch_floats(B, C)), dtype=torch.float32, device=self.device)
        self.keep.append(scratch)
        xv, yv = rt.view(x_fast), rt.view(y_slice)
        self.keep += [xv, yv]
        k = int(w.numel())
        self._add(lambda s: rt.check(
            L.esf_eca_fuse(ctypes.byref(xv), alpha, w.data_ptr(), k, sc.data_ptr(), sh.data_ptr(), scratch.data_ptr(),
                           ctypes.byref(yv), s), "esf_eca_fuse"), "eca_fuse", "C=%d" % C,
            nbytes=2 * self._nbytes(x_fast) + self._nbytes(y_slice), launches=2)

    def position_attention(self, x_slow, y_slice, alpha, w_down, att, bn):
        """1x1x1 C->d conv composed with the q/k/v 1x1x1 convs into one GEMM (FP32 out), hi/lo packing, fused
        flash-style attention + gamma*O + x + BN + ReLU + x alpha upsample + concat
        (custom_video_model_builder.py:141-146, wdf_attention_helper.py:33-54)."""
        B, T, H, W, C = x_slow.shape
        wd = host64(w_down).reshape(w_down.shape[0], C)                      # (d, C)
        d = wd.shape[0]
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
        proj = self.act(B, T, H, W, 4 * d, dtype=torch.float32)
        self.conv(x_slow, proj, w_all, b_all)
        L = rt.lib()
        N = T * H * W
        if d > 128:
            scale, shift = bn_affine(bn)
            sc, sh = self.tensor(scale), self.tensor(shift)
            gamma = float(att.gamma.detach().float().item())
            yv = rt.view(y_slice)
            self.keep.append(yv)
            self._add(lambda s: rt.check(
                L.esf_attn_generic(proj.data_ptr(), B, T, H, W, d, gamma, sc.data_ptr(), sh.data_ptr(), alpha,
                                   ctypes.byref(yv), s), "esf_attn_generic"), "attention", "N=%d d=%d generic" % (N, d),
                flops=4.0 * B * N * N * d, exps=float(B) * N * N, nbytes=self._nbytes(proj, y_slice))
            return
        if self.attn_impl == "tcgen05":
            nbytes = int(L.esf_attn_tc_pack_bytes(B, N, d))
            if nbytes < 0:
                rt.check(nbytes, "esf_attn_tc_pack_bytes")
            packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.keep.append(packed)
            scale, shift = bn_affine(bn)
            sc, sh = self.tensor(scale), self.tensor(shift)
            gamma = float(att.gamma.detach().float().item())
            yv = rt.view(y_slice)
            h = ctypes.c_void_p()
            rt.check(L.esf_attn_tc_create(packed.data_ptr(), B, T, H, W, d, gamma, sc.data_ptr(), sh.data_ptr(), alpha,
                                          ctypes.byref(yv), ctypes.byref(h)), "esf_attn_tc_create")
            self.handles.append(h)
            self._add(lambda s: rt.check(L.esf_attn_tc_pack(proj.data_ptr(), B, N, d, self.a16, packed.data_ptr(), s),
                                         "esf_attn_tc_pack"), "attn_pack", "N=%d d=%d" % (N, d),
                      nbytes=self._nbytes(proj) + nbytes)
            self._add(lambda s, h=h: rt.check(L.esf_op_launch(h, s), "esf_op_launch"), "attention",
                      "N=%d d=%d" % (N, d), flops=4.0 * B * N * N * d, exps=float(B) * N * N,
                      nbytes=nbytes + self._nbytes(y_slice))
            return
        nbytes = int(L.esf_attn_pack_bytes(B, N, d))
        if nbytes < 0:
            rt.check(nbytes, "esf_attn_pack_bytes")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.keep.append(packed)
        scale, shift = bn_affine(bn)
        sc, sh = self.tensor(scale), self.tensor(shift)
        gamma = float(att.gamma.detach().float().item())
        yv = rt.view(y_slice)
        self.keep.append(yv)
        self._add(lambda s: rt.check(L.esf_attn_pack(proj.data_ptr(), B, N, d, packed.data_ptr(), s),
                                     "esf_attn_pack"), "attn_pack", "N=%d d=%d" % (N, d),
                  nbytes=self._nbytes(proj) + nbytes)
        self._add(lambda s: rt.check(
            L.esf_attn_fused(packed.data_ptr(), B, T, H, W, d, gamma, sc.data_ptr(), sh.data_ptr(), alpha,
                             ctypes.byref(yv), s), "esf_attn_fused"), "attention", "N=%d d=%d" % (N, d),
            flops=4.0 * B * N * N * d, exps=float(B) * N * N, nbytes=nbytes + self._nbytes(y_slice))

    def head(self, xs, weight, bias, act):
        B = xs[0].shape[0]
        cin = sum(x.shape[4] for x in xs)
        K = weight.shape[0]
        feat = torch.empty((B, cin), dtype=torch.float32, device=self.device)
        out = torch.empty((B, K), dtype=torch.float32, device=self.device)
        w, b = self.tensor(weight), self.tensor(bias)
        v0 = rt.view(xs[0])
        v1 = rt.view(xs[1]) if len(xs) > 1 else rt.null_view()
        self.keep += [feat, out, v0, v1]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_head_pool(ctypes.byref(v0), ctypes.byref(v1), feat.data_ptr(), s),
                                     "esf_head_pool"), "head_pool", "", nbytes=self._nbytes(*xs), launches=len(xs))
        self._add(lambda s: rt.check(
            L.esf_head_fc(feat.data_ptr(), B, cin, cin, w.data_ptr(), b.data_ptr(), K, act, out.data_ptr(), K, s),
            "esf_head_fc"), "head_fc", "", flops=2.0 * B * cin * K, nbytes=self._nbytes(feat, w, out),
            launches=head_fc_launches(B, cin, K, act))
        self.out = out
        return out

    def head_positions(self, xs, pool_sizes, weight, bias, act):
        """ResNetBasicHead.forward, eval branch, general case (head_helper.py:198-223): per-pathway AvgPool3d(kernel,
        stride 1) leaves P = T' * H' * W' positions; Linear + activation at every position, mean over the positions."""
        B = xs[0].shape[0]
        pooled = []
        for x, ps in zip(xs, pool_sizes):
            _, T, H, W, C = x.shape
            kt, kh, kw = [int(v) for v in ps]
            y = self.act(B, T - kt + 1, H - kh + 1, W - kw + 1, C)
            self.pool(x, y, (kt, kh, kw), (1, 1, 1), (0, 0, 0), is_avg=True)
            pooled.append(y)
        pos = tuple(pooled[0].shape[1:4])
        assert all(tuple(y.shape[1:4]) == pos for y in pooled), "pathway dimensions are not consistent."
        P = pos[0] * pos[1] * pos[2]
        rows = []
        for y in pooled:    # (B, T', H', W', C) -> one 1x1x1 "clip" per (clip, position)
            C = y.shape[4]
            assert y.stride(3) * pos[2] == y.stride(2) and y.stride(2) * pos[1] == y.stride(1) and y.stride(1) * pos[0] == y.stride(0)
            rows.append(y.as_strided((B * P, 1, 1, 1, C), (y.stride(3), y.stride(3), y.stride(3), y.stride(3), 1),
                                     y.storage_offset()))
        per_pos = self.head(rows, weight, bias, act)
        K = weight.shape[0]
        out = torch.empty((B, K), dtype=torch.float32, device=self.device)
        self.keep += [out]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_group_mean(per_pos.data_ptr(), B, P, K, out.data_ptr(), s), "esf_group_mean"),
                  "head_mean", "P=%d" % P, nbytes=self._nbytes(per_pos, out))
        self.out = out
        return out

    def pooled_fc(self, x, weight, bias, act, out, out_stride):
        """global mean over (T,H,W) of one activation -> FC (+bias, act) -> `out` (FP32 rows of stride out_stride):
        the `avg_pool3d -> conv_head_{slow,fast} -> ReLU` tail of GhostNetBasicHead (head_helper.py:676-688)."""
        B, C = x.shape[0], x.shape[4]
        K = weight.shape[0]
        feat = torch.empty((B, C), dtype=torch.float32, device=self.device)
        w, b = self.tensor(weight.reshape(K, C)), self.tensor(bias)
        xv, nv = rt.view(x), rt.null_view()
        self.keep += [feat, xv, nv, out]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_head_pool(ctypes.byref(xv), ctypes.byref(nv), feat.data_ptr(), s),
                                     "esf_head_pool"), "head_pool", "", nbytes=self._nbytes(x))
        self._add(lambda s: rt.check(L.esf_head_fc(feat.data_ptr(), B, C, C, w.data_ptr(), b.data_ptr(), K, act,
                                                   out.data_ptr(), out_stride, s), "esf_head_fc"), "head_fc", "",
                  launches=head_fc_launches(B, C, K, act))

    def fc(self, feat, weight, bias, act):
        """out = act(feat @ weight^T + bias) on FP32 rows (final classifier of the GhostNet head)."""
        B, cin = feat.shape
        K = weight.shape[0]
        out = torch.empty((B, K), dtype=torch.float32, device=self.device)
        w, b = self.tensor(weight), self.tensor(bias)
        self.keep += [out, feat]
        L = rt.lib()
        self._add(lambda s: rt.check(L.esf_head_fc(feat.data_ptr(), B, cin, cin, w.data_ptr(), b.data_ptr(), K, act,
                                                   out.data_ptr(), K, s), "esf_head_fc"), "head_fc", "",
                  launches=head_fc_launches(B, cin, K, act))
        self.out = out
        return out

    # ---------------------------------------------------------------- execution
    def launch_all(self):
        s = rt.current_stream_ptr()
        for op in self.ops:
            op(s)

    def launch_graph_ops(self):
        s = rt.current_stream_ptr()
        for op, eager in zip(self.ops, self.eager):
            if not eager:
                op(s)

    def launch_eager_ops(self):
        s = rt.current_stream_ptr()
        for op, eager in zip(self.ops, self.eager):
            if eager:
                op(s)

    def capture(self):
        """Record every launch that only touches plan-owned memory into one CUDA graph (replayed by run()); the ops
        that read the caller's clips (stem packing) stay outside and run ahead of the replay."""
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.launch_all()  # warm-up outside capture: sets kernel attributes, faults in the code
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch_graph_ops()
        self.graph = g

    def profile_ops(self, repeats=3):
        """Per-op device time (ms, best of `repeats`) with CUDA events on the launching stream, eager launches."""
        s = rt.current_stream_ptr()
        times = [float("inf")] * len(self.ops)
        for _ in range(repeats):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.ops) + 1)]
            evs[0].record()
            for i, op in enumerate(self.ops):
                op(s)
                evs[i + 1].record()
            torch.cuda.synchronize(self.device)
            for i in range(len(self.ops)):
                times[i] = min(times[i], evs[i].elapsed_time(evs[i + 1]))
        return times

    def run_frames(self, fill):
        """Frame route: `fill()` launches the kernels that write the stem inputs from uint8 frames (frames.py); the
        FP32 stem-pack launches are skipped, every other launch is the same as in run()."""
        self.input_override = {}
        fill()
        s = rt.current_stream_ptr()
        for op, eager, meta in zip(self.ops, self.eager, self.meta):
            if meta.get("is_pack"):
                continue
            if eager or self.graph is None:
                op(s)
        if self.graph is not None:
            self.graph.replay()
        return self.out

    def run(self, inputs=None, gather=None):
        """`inputs`: optional caller tensors to read INSTEAD of the plan-owned input buffers (same shape, FP32,
        contiguous, same device; None entries keep the plan's buffer) -- saves the device-to-device staging copy.
        `gather`: {pathway: (source clip tensor, device int32 frame index)} -- that pathway's banded stem packs its rows
        out of the source clip's frames (Plan.stem / esf_stem_pack_gather)."""
        self.input_override = {}
        if inputs is not None:
            self.input_override = {own.data_ptr(): src.data_ptr() for own, src in zip(self.inputs, inputs)
                                   if src is not None}
        self.gather = {}
        for pw, (src, idx) in (gather or {}).items():
            assert self.inputs[pw].data_ptr() in self.stem_routes
            self.gather[self.inputs[pw].data_ptr()] = (src.data_ptr(), src.shape[2], idx.data_ptr())
        try:
            if self.graph is not None:
                self.launch_eager_ops()
                self.graph.replay()
            else:
                self.launch_all()
        finally:
            self.input_override = {}
            self.gather = {}
        return self.out

    def __del__(self):
        try:
            L = rt.lib()
            for h in self.handles:
                L.esf_op_destroy(h)
        except Exception:
            pass
