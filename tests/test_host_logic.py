"""Host-side logic that needs no GPU: BN folding, implicit-GEMM weight packing, registry / build_model / cfg."""
import pytest
import torch
import torch.nn.functional as F

import efficient_slowfast_b200 as esf
from efficient_slowfast_b200 import engine, runtime as rt


def test_fold_conv_bn_equals_conv_then_bn():
    g = torch.Generator().manual_seed(0)
    conv = torch.nn.Conv3d(6, 10, (3, 1, 1), padding=(1, 0, 0), bias=True)
    bn = torch.nn.BatchNorm3d(10)
    bn.weight.data = torch.rand(10, generator=g) + 0.5
    bn.bias.data = torch.randn(10, generator=g)
    bn.running_mean = torch.randn(10, generator=g)
    bn.running_var = torch.rand(10, generator=g) + 0.1
    bn.eval()
    x = torch.randn(2, 6, 5, 4, 4, generator=g)
    with torch.no_grad():
        ref = bn(conv(x))
    w, b = engine.fold_conv_bn(conv.weight, conv.bias, bn)
    got = F.conv3d(x.double(), w, b, padding=(1, 0, 0))
    assert torch.allclose(got.float(), ref, atol=1e-5, rtol=1e-5)


def test_pack_igemm_weight_layout(esf_lib):
    cout, cin, k = 20, 72, (1, 3, 3)
    w = torch.arange(cout * cin * 9, dtype=torch.float64).reshape(cout, cin, *k) / 1000.0
    b = torch.arange(cout, dtype=torch.float64)
    wp, bp = engine.pack_igemm_weight(w, b, "cpu")
    kc, kchunks, n_tile, n_pad = rt.igemm_geometry(cin, cout)
    assert wp.shape == (n_pad, 9 * kchunks * kc) and wp.dtype == torch.bfloat16
    wp = wp.float().reshape(n_pad, 9, kchunks * kc)
    for tap, (kh, kw) in enumerate([(i, j) for i in range(3) for j in range(3)]):
        assert torch.equal(wp[:cout, tap, :cin], w[:, :, 0, kh, kw].bfloat16().float())
    assert (wp[cout:] == 0).all() and (wp[:, :, cin:] == 0).all()
    assert torch.equal(bp[:cout], b.float()) and (bp[cout:] == 0).all()


def test_registry_and_build_model():
    names = esf.MODEL_REGISTRY.names()
    assert "SlowFast" in names and "SlowFastDualAttention" in names
    with pytest.raises(KeyError):
        esf.MODEL_REGISTRY.get("NoSuchModel")
    with pytest.raises(AssertionError):
        @esf.MODEL_REGISTRY.register()
        class SlowFast:  # duplicate name
            pass
    cfg = esf.slowfast_4x16_r50_cfg()
    cfg.NUM_GPUS = 0
    m = esf.build_model(cfg)
    assert isinstance(m, torch.nn.Module)
    assert sum(p.numel() for p in m.parameters()) == 34480216   # SlowFast 4x16 R50 (34.48 M)
    cfg.NUM_GPUS = 10 ** 6
    with pytest.raises(AssertionError):
        esf.build_model(cfg)


def test_cfg_merge_semantics(tmp_path):
    cfg = esf.get_cfg()
    p = tmp_path / "x.yaml"
    p.write_text("SLOWFAST:\n  ALPHA: 4\nMODEL:\n  MODEL_NAME: SlowFastDualAttention\n")
    cfg.merge_from_file(str(p))
    cfg.merge_from_list(["DATA.NUM_FRAMES", "32", "RESNET.SPATIAL_STRIDES", "[[1, 1], [2, 2], [2, 2], [2, 2]]"])
    assert cfg.SLOWFAST.ALPHA == 4 and cfg.SLOWFAST.BETA_INV == 8
    assert cfg.DATA.NUM_FRAMES == 32 and cfg.RESNET.SPATIAL_STRIDES[1] == [2, 2]
    c2 = cfg.clone()
    c2.SLOWFAST.ALPHA = 8
    assert cfg.SLOWFAST.ALPHA == 4


def test_zero_init_final_bn_and_gamma_zero():
    cfg = esf.slowfast_dual_8x8_r50_cfg()
    cfg.NUM_GPUS = 0
    m = esf.build_model(cfg)
    assert float(m.s2.pathway0_res0.branch2.c_bn.weight.abs().sum()) == 0.0
    assert float(m.s2.pathway0_res0.branch2.a_bn.weight.min()) == 1.0
    assert float(m.s1_fuse.attention_spatial_s2f.gamma) == 0.0
    assert m.s1_fuse.attention_channel_f2s.conv.weight.shape == (1, 1, 3)


def _banded_gemm_reference(x_ndhwc, band, bias_t, cout, k, stride_hw, pad, WB):
    """CPU emulation of what esf_conv_wfold_create computes: for every block of WB output columns, the GEMM row is the
    contiguous run of ((WB-1)*sW + kW) input columns x C, zero outside the tensor; K = taps(kt,kh) x padded window."""
    B, T, H, W, C = x_ndhwc.shape
    kt, kh, kw = k
    sH, sW = stride_hw
    pT, pH, pW = pad
    To, Ho, Wo = T + 2 * pT - kt + 1, (H + 2 * pH - kh) // sH + 1, (W + 2 * pW - kw) // sW + 1
    win = ((WB - 1) * sW + kw) * C
    kpad = -(-win // 64) * 64
    band = band.double().reshape(band.shape[0], kt * kh, kpad)
    xp = F.pad(x_ndhwc.double(), (0, 0, pW, pW + kpad, pH, pH, pT, pT))   # generous right padding = TMA zero fill
    out = torch.zeros(B, To, Ho, Wo, cout, dtype=torch.float64)
    for cb in range(Wo // WB):
        w0 = cb * WB * sW
        for t in range(To):
            for h in range(Ho):
                acc = bias_t[:WB * cout].double().clone().repeat(B, 1)
                for it in range(kt):
                    for ih in range(kh):
                        row = xp[:, t + it, h * sH + ih, w0:w0 + kpad // C + 1].reshape(B, -1)[:, :kpad]
                        acc += row @ band[:WB * cout, it * kh + ih].T
                out[:, t, h, cb * WB:(cb + 1) * WB] = acc.reshape(B, WB, cout)
    return out


@pytest.mark.parametrize("cin,cout,k,s,WB", [(8, 8, (1, 3, 3), (1, 1), 4), (16, 8, (3, 1, 1), (1, 1), 8),
                                             (8, 32, (1, 1, 1), (1, 1), 8), (16, 16, (1, 3, 3), (2, 2), 2)])
def test_pack_wfold_band_is_the_convolution(esf_lib, cin, cout, k, s, WB):
    g = torch.Generator().manual_seed(cin + cout)
    pad = (k[0] // 2, k[1] // 2, k[2] // 2)
    x = torch.randn(2, 3, 6, 8 * s[1], cin, generator=g)
    w = torch.randn(cout, cin, *k, generator=g).double()
    b = torch.randn(cout, generator=g).double()
    band, bt = engine.pack_wfold_band(w, b, WB, s[1], "cpu", torch.float32)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w, b, (1, s[0], s[1]), pad).permute(0, 2, 3, 4, 1)
    got = _banded_gemm_reference(x, band, bt, cout, k, s, pad, WB)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


def _tband_reference(x_cl, band, bias_t, cout, k, s, pad, WB, nslots):
    """CPU model of stem_tband_kernel's schedule: input frames in order, one GEMM per (g, kh) against the weights of all
    time taps, output frames accumulating in a ring of nslots slots (first touch overwrites, last touch publishes)."""
    B, T, H, W, C = x_cl.shape
    kt, kh, kw = k
    sH, sW = s
    pT, pH, pW = pad
    To, Ho, Wo = T + 2 * pT - kt + 1, (H + 2 * pH - kh) // sH + 1, (W + 2 * pW - kw) // sW + 1
    NB = WB * cout
    xp = F.pad(x_cl.double(), (0, 0, pW, pW + 64, pH, pH + sH))           # zero rows / columns the TMA box would fill
    band = band.double().reshape(kt * NB, kh, 64)
    out = torch.full((B, To, Ho, Wo, cout), float("nan"), dtype=torch.float64)
    for cb in range(Wo // WB):
        w0 = cb * WB * sW
        slots = torch.full((nslots, B, Ho, NB), float("nan"), dtype=torch.float64)
        for g in range(T):
            t_base = g + pT - kt + 1
            t_lo, t_hi = max(t_base, 0), min(g + pT, To - 1)
            new_lo = 0 if g == 0 else g + pT
            for ih in range(kh):
                rows = xp[:, g, ih:ih + sH * Ho:sH, w0:w0 + 64 // C + 2].reshape(B, Ho, -1)[:, :, :64]
                for t in range(t_lo, t_hi + 1):
                    u = t - t_base
                    d = rows @ band[u * NB:(u + 1) * NB, ih].T
                    if ih == 0 and t >= new_lo:
                        slots[t % nslots] = d
                    else:
                        slots[t % nslots] += d
            for t in range(t_lo, t_hi + 1):
                if g == min(t + kt - 1 - pT, T - 1):
                    out[:, t, :, cb * WB:(cb + 1) * WB] = (slots[t % nslots] + bias_t.double()).reshape(B, Ho, WB, cout)
                    slots[t % nslots] = float("nan")
    return out


@pytest.mark.parametrize("cin,cout,k,s,WB,T", [(3, 8, (5, 7, 7), (2, 2), 4, 11), (1, 8, (5, 7, 7), (2, 2), 4, 6),
                                               (3, 4, (3, 3, 3), (1, 1), 4, 7), (3, 8, (5, 7, 7), (2, 2), 4, 2)])
def test_pack_stem_tband_is_the_convolution(esf_lib, cin, cout, k, s, WB, T):
    g = torch.Generator().manual_seed(cin + cout + T)
    pad = (k[0] // 2, k[1] // 2, k[2] // 2)
    x = torch.randn(2, T, 10, 8 * s[1], cin, generator=g)
    w = torch.randn(cout, cin, *k, generator=g).double()
    b = torch.randn(cout, generator=g).double()
    band, bt = engine.pack_stem_tband(w, b, WB, s[1], "cpu", torch.float32)
    assert band.shape == (k[0] * WB * cout, k[1] * 64) and bt.shape == (WB * cout,)
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3).double(), w, b, (1, s[0], s[1]), pad).permute(0, 2, 3, 4, 1)
    got = _tband_reference(x, band, bt, cout, k, s, pad, WB, nslots=4 if k[0] == 3 else 8)
    assert got.shape == ref.shape and not torch.isnan(got).any()
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)


def test_stem_tband_planner(esf_lib):
    """esf_stem_tband_wb (pure host code): which stems take the temporal-band kernel.  The fast pathway's 5x7x7 3->8 and
    its grey-scale form do (WB = 4: N = 5 x 32 <= 256, 16 TMEM slots); kT = 1 stems, wide stems and odd widths do not."""
    wb = rt.stem_tband_wb
    assert wb(224, 3, 8, 5, 7, 7, 2, 3) == 4
    assert wb(224, 1, 8, 5, 7, 7, 2, 3) == 4
    assert wb(64, 3, 16, 3, 7, 7, 2, 3) == 4          # 64-column slots, 8 of them
    assert wb(224, 3, 64, 1, 7, 7, 2, 3) == 0          # no time taps: the banded stem
    assert wb(224, 3, 64, 5, 7, 7, 2, 3) == 0          # kT x WB x Cout > 256 columns for every block width
    assert wb(224, 3, 24, 3, 3, 3, 2, 1) == 0          # WB x Cout not a power-of-two slot width
    assert wb(220, 3, 8, 5, 7, 7, 2, 3) in (0, 2)      # 110 output columns: not a multiple of 4
    assert wb(0, 3, 8, 5, 7, 7, 2, 3) == 0


def test_wfold_block_planner():
    def views(B, T, H, W, cin, cout, slice_out=False):
        x = torch.empty(B, T, H, W, cin)
        yb = torch.empty(B, T, H, W, cout + (16 if slice_out else 0))
        return x, (yb[..., 8:8 + cout] if slice_out else yb)
    x, y = views(1, 4, 56, 56, 8, 8)
    assert engine.wfold_block(x, y, None, (8, 8, 1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)) == 14
    x, y = views(1, 4, 56, 56, 8, 32, slice_out=True)          # sliced destination: WB * Cout must be a power of two
    assert engine.wfold_block(x, y, None, (32, 8, 1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1)) == 8
    x, y = views(1, 4, 14, 14, 64, 64)                          # C_in > 32: plain implicit GEMM
    assert engine.wfold_block(x, y, None, (64, 64, 1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)) == 0
    x, y = views(1, 4, 56, 56, 8, 16)                           # temporal stride: not folded
    assert engine.wfold_block(x, y, None, (16, 8, 5, 1, 1), (4, 1, 1), (2, 0, 0), (1, 1, 1)) == 0
    xs = torch.empty(1, 4, 56, 56, 24)[..., :8]                 # input is a channel slice (not dense in W, C)
    assert engine.wfold_block(xs, y, None, (16, 8, 1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1)) == 0


def test_activation_pitch_rule():
    plan = engine.Plan("cpu", "fp16")
    for C, pitch in [(3, 3), (7, 7), (8, 8), (27, 32), (180, 184), (540, 544), (64, 64)]:
        t = plan.act(1, 2, 3, 4, C)
        assert t.shape[4] == C and t.stride(3) == pitch and t.dtype == torch.float16
        assert engine.Plan._aligned(t) == (C >= 8)
    assert plan.act(1, 2, 3, 4, 12, dtype=torch.float32).stride(3) == 12     # FP32 projections stay dense


def test_head_fc_launch_accounting():
    assert engine.head_fc_launches(64, 2304, 400, rt.HEAD_SOFTMAX) == 2
    assert engine.head_fc_launches(64, 2304, 400, rt.HEAD_RELU) == 1
    assert engine.head_fc_launches(2, 2304, 400, rt.HEAD_SOFTMAX) == 1
    assert engine.head_fc_launches(2, 20000, 400, rt.HEAD_SOFTMAX) == 2


# ------------------------------------------------------------------------------------------------ Caffe2 checkpoints
def test_caffe2_name_mapping_matches_reference_fixture():
    """tests/golden/caffe2_names.json: 646 Caffe2 blob names (stems, bottlenecks of both pathways, lateral convs,
    Nonlocal blocks, classifier) converted by the reference's utils/c2_model_loading.get_name_convert_func."""
    import json
    import os

    import helpers
    from efficient_slowfast_b200.checkpoint import caffe2_to_pytorch_name

    want = json.load(open(os.path.join(helpers.GOLDEN_DIR, "caffe2_names.json")))
    assert len(want) > 600
    for c2, key in want.items():
        assert caffe2_to_pytorch_name(c2) == key, c2


def test_caffe2_checkpoint_loads(tmp_path):
    """A Caffe2-format pickle ({"blobs": {name: ndarray}}, checkpoint.py:206-259) fills the drop-in model."""
    import json
    import os
    import pickle

    import numpy as np
    import torch

    import efficient_slowfast_b200 as esf
    import helpers

    cfg = esf.resnet_cfg("slow", nln=True)
    cfg.NUM_GPUS = 0
    model = esf.build_model(cfg)
    sd = model.state_dict()
    names = json.load(open(os.path.join(helpers.GOLDEN_DIR, "caffe2_names.json")))
    rng = np.random.RandomState(0)
    blobs, used = {}, set()
    for c2, key in sorted(names.items()):
        if key in sd and key not in used:       # 'conv1_w' and 'res_conv1_w' name the same tensor
            used.add(key)
            blobs[c2] = rng.randn(*sd[key].shape).astype(np.float32)
    names = {c2: names[c2] for c2 in blobs}
    assert len(blobs) > 100
    blobs["lr"] = np.zeros(1, np.float32)
    blobs["res2_0_branch2a_w_momentum"] = np.zeros(3, np.float32)
    blobs["res9_0_branch2a_w"] = np.zeros(3, np.float32)            # no such layer: reported, not loaded
    path = str(tmp_path / "c2.pkl")
    with open(path, "wb") as f:
        pickle.dump({"blobs": blobs}, f)
    assert esf.load_checkpoint(path, model, False, convert_from_caffe2=True) == -1
    after = model.state_dict()
    for c2, key in names.items():
        if key in sd:
            assert torch.equal(after[key], torch.tensor(blobs[c2])), key
    assert esf.load_checkpoint.last_report["skipped"] == ["res9_0_branch2a_w"]


def test_fp32_path_operand_split_reproduces_fp32_conv():
    """The algebra of the FP32-accurate plan (engine_fp32.py), on the host: activation planes [hi | lo | hi], weight
    [w_hi | w_hi | w_lo] with power-of-two row scaling -> an ordinary convolution over 3 C channels whose result,
    un-scaled, equals the FP32 convolution to ~2^-22 -- also when the input is a channel slice of a wider concat
    buffer (zero weights on the other channels)."""
    import torch.nn.functional as F
    from efficient_slowfast_b200.engine_fp32 import split_weight_rows

    g = torch.Generator().manual_seed(0)
    ctot, c0, cin, cout = 24, 8, 16, 12
    x = torch.randn(2, ctot, 3, 5, 5, generator=g) * 3.0
    w = torch.randn(cout, cin, 3, 1, 1, generator=g).double() * 0.02 * torch.logspace(-3, 1, cout).double().view(-1, 1, 1, 1, 1)
    hi, lo, inv = split_weight_rows(w)
    assert torch.equal(hi, hi.half().double()) and torch.equal(lo, lo.half().double())
    assert (lo.abs().amax(dim=(1, 2, 3, 4)) > 6.2e-5).all()          # w_lo stays a normal FP16 number in every row
    plane = ctot
    w3 = torch.zeros(cout, 3 * plane, 3, 1, 1, dtype=torch.float64)
    w3[:, c0:c0 + cin] = hi
    w3[:, plane + c0:plane + c0 + cin] = hi
    w3[:, 2 * plane + c0:2 * plane + c0 + cin] = lo
    x_hi = x.half()
    x_lo = (x - x_hi.float()).half()
    x3 = torch.cat([x_hi, x_lo, x_hi], dim=1).double()
    got = F.conv3d(x3, w3, padding=(1, 0, 0)) * inv.view(1, -1, 1, 1, 1)
    want = F.conv3d(x[:, c0:c0 + cin].double(), w, padding=(1, 0, 0))
    err = ((got - want).abs().amax(dim=(0, 2, 3, 4)) / want.abs().amax(dim=(0, 2, 3, 4))).max().item()
    assert err < 2e-6, err
    # the same product with single FP16 operands is three orders of magnitude worse
    plain = F.conv3d(x[:, c0:c0 + cin].half().double(), w.half().double(), padding=(1, 0, 0))
    assert ((plain - want).abs().amax() / want.abs().amax()).item() > 1e-4
