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
