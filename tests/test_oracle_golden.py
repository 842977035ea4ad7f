"""Pins the CPU oracle (oracle/slowfast_oracle.py) to golden vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

import helpers
import recipe
from oracle import slowfast_oracle as O


@pytest.mark.parametrize("name,tag", [("dual_r50", "s64"), ("slowfast_r50", "s64"), ("dual_r50", "s224"),
                                      ("slowfast_r50", "s224"), ("slowfast_r50_stress", "s64"),
                                      ("dual_r50_stress", "s64"), ("shufflenetv2_w05", "s112"),
                                      ("shufflenetv2_w05", "s224"), ("shufflenet_w2g3", "s112"),
                                      ("shufflenet_w2g3", "s64"), ("mobilenetv2_w1", "s112"),
                                      ("mobilenetv2_w1", "s224"), ("ghostnet_w1", "s112"), ("ghostnet_w1", "s64"), ("ghostnet_w1", "s224"),
                                      ("i3d_r50", "s224"), ("slow_r50", "s64"),
                                      ("slow_nln_r50", "s64"), ("i3d_nln_r50", "s96"),
                                      ("slowfast_r50_fcn", "s96"), ("slowfast_r50_fcn", "s64"), ("slow_r50", "s96"),
                                      ("slowfast_r50_g2", "s64"), ("dual_r18_gray", "s112"), ("dual_r18_gray", "s128"),
                                      ("fast_r18_gray", "s112"), ("slowfast_r101", "s64"), ("slowfast_r50_sigmoid", "s64")])
def test_oracle_matches_reference_golden(name, tag):
    cfg, model, gold = helpers.case_model_and_weights(name)
    xs = helpers.case_inputs(name, tag)
    taps = {}
    y = O.forward(cfg, model.state_dict(), xs, taps=taps)
    # Not bit-exact by construction: FP32 CPU kernels re-associate sums differently on another host, and the
    # row-chunked softmax over N = 25 088 keys sums in a different order than the reference's materialised N x N
    # form (measured FP32 noise floor of this model: ~2e-4, SURVEY.md finding 8).
    tol = 1e-3 if tag == "s224" and name != "slowfast_r50" else 1e-4
    assert helpers.rel_err(y, gold[tag + "/probs"]) < tol
    assert helpers.rel_err(taps["logits"].reshape(y.shape[0], -1), gold[tag + "/logits"]) < tol
    assert torch.equal(y.argmax(1), torch.as_tensor(gold[tag + "/probs"]).argmax(1))
    for sname in ("s1", "s1_fuse", "s2", "s2_fuse", "s3", "s3_fuse", "s4", "s4_fuse", "s5", "s5_fuse", "s6", "s7",
                  "s7_fuse", "s8"):
        if sname not in taps:
            continue            # the ShuffleNet models end at s4_fuse
        for pw in range(len(taps[sname])):
            t = taps[sname][pw]
            assert list(t.shape) == list(gold["%s/%s/%d/shape" % (tag, sname, pw)])
            flat = t.reshape(-1)
            smp = flat[recipe.sample_indices(flat.numel())]
            ref = gold["%s/%s/%d/samples" % (tag, sname, pw)]
            scale = gold["%s/%s/%d/stats" % (tag, sname, pw)][2]
            assert np.abs(smp.numpy() - ref).max() <= tol * scale, (sname, pw)
    if name != "ghostnet_w1" and cfg.MODEL.HEAD_ACT == "softmax":   # GhostNet: ReLU(logits); sigmoid heads: multi-label
        assert abs(y.sum(1) - 1).max() < 1e-5


def test_oracle_default_init_corner():
    """gamma = 0 and zero-initialised final BN (the reference's shipped init): attention/bottleneck branches vanish."""
    import efficient_slowfast_b200 as esf

    gold = helpers.load_golden("default_init")
    for name in ("dual_r50", "slowfast_r50"):
        cfg = helpers.case_cfg(name)
        torch.manual_seed(1234)
        model = esf.build_model(cfg).eval()   # same module order + same init ops as the reference => same weights
        xs = recipe.pack_pathway_output(recipe.seeded_clip(2, 32, 64, seed=1), cfg.SLOWFAST.ALPHA)
        y = O.forward(cfg, model.state_dict(), xs)
        assert helpers.rel_err(y, gold[name + "/probs"]) < 1e-4


def test_row_chunked_attention_equals_materialised():
    """The query-row-chunked restatement is the same arithmetic as the reference's N x N softmax(bmm)."""
    g = torch.Generator().manual_seed(3)
    d, T, H, W = 8, 2, 6, 5
    sd = {"a.gamma": torch.tensor([0.7])}
    for n in ("query_conv", "key_conv", "value_conv"):
        sd["a.%s.weight" % n] = torch.randn(d, d, 1, 1, 1, generator=g)
        sd["a.%s.bias" % n] = torch.randn(d, generator=g)
    x = torch.randn(2, d, T, H, W, generator=g)
    full = O.position_attention(x, sd, "a", row_chunk=10 ** 9)
    chunked = O.position_attention(x, sd, "a", row_chunk=7)
    assert torch.allclose(full, chunked, atol=1e-6, rtol=1e-6)
    # explicit N x N form of wdf_attention_helper.py:42-53
    N = T * H * W
    q = torch.nn.functional.conv3d(x, sd["a.query_conv.weight"], sd["a.query_conv.bias"]).view(2, -1, N).permute(0, 2, 1)
    k = torch.nn.functional.conv3d(x, sd["a.key_conv.weight"], sd["a.key_conv.bias"]).view(2, -1, N)
    v = torch.nn.functional.conv3d(x, sd["a.value_conv.weight"], sd["a.value_conv.bias"]).view(2, -1, N)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    out = 0.7 * torch.bmm(v, att.permute(0, 2, 1)).view(2, d, T, H, W) + x
    assert torch.allclose(full, out, atol=1e-6, rtol=1e-6)
