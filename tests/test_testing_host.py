"""Host-side callers of the forward path: view ensembling (TestMeter), top-k, checkpoint wire format."""
import pytest
import torch

import efficient_slowfast_b200 as esf
from oracle import ref_shim


def _loop_meter(num_videos, num_clips, num_cls, method, batches):
    """Independent restatement of the per-clip update loop (meters.py:279-310) used as the checker."""
    preds_v = torch.zeros(num_videos, num_cls)
    labels_v = torch.zeros(num_videos).long()
    count = torch.zeros(num_videos).long()
    for preds, labels, ids in batches:
        for i in range(preds.shape[0]):
            v = int(ids[i]) // num_clips
            labels_v[v] = labels[i]
            preds_v[v] = preds_v[v] + preds[i] if method == "sum" else torch.max(preds_v[v], preds[i])
            count[v] += 1
    return preds_v, labels_v, count


def _batches(num_videos, num_clips, num_cls, bs, seed):
    g = torch.Generator().manual_seed(seed)
    labels_v = torch.randint(1, num_cls, (num_videos,), generator=g)
    ids = torch.randperm(num_videos * num_clips, generator=g)
    out = []
    for i in range(0, len(ids), bs):
        c = ids[i:i + bs]
        out.append((torch.rand(len(c), num_cls, generator=g), labels_v[c // num_clips], c))
    return out


@pytest.mark.parametrize("method", ["sum", "max"])
def test_meter_matches_per_clip_loop(method):
    nv, nc, ncls = 13, 6, 17
    batches = _batches(nv, nc, ncls, 8, seed=3)      # batches of 8 hold several clips of the same video
    m = esf.TestMeter(nv, nc, ncls, len(batches), ensemble_method=method)
    for b in batches:
        m.update_stats(*b)
    p, l, c = _loop_meter(nv, nc, ncls, method, batches)
    assert torch.allclose(m.video_preds, p, atol=1e-6) and torch.equal(m.video_labels, l) and torch.equal(m.clip_count, c)
    stats = m.finalize_metrics(ks=(1, 5))
    top = torch.topk(p, 5, dim=1).indices
    assert stats["top1_acc"] == "{:.2f}".format(float((top[:, 0] == l).float().mean()) * 100)
    assert stats["top5_acc"] == "{:.2f}".format(float((top == l[:, None]).any(1).float().mean()) * 100)
    assert stats["complete"]
    m.reset()
    assert m.video_preds.abs().sum() == 0 and m.clip_count.sum() == 0


def test_meter_rejects_conflicting_labels_and_bad_method():
    m = esf.TestMeter(2, 2, 3, 1)
    m.update_stats(torch.rand(1, 3), torch.tensor([2]), torch.tensor([0]))
    with pytest.raises(AssertionError):
        m.update_stats(torch.rand(1, 3), torch.tensor([1]), torch.tensor([1]))
    with pytest.raises(NotImplementedError):
        esf.TestMeter(2, 2, 3, 1, ensemble_method="mean")


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("method", ["sum", "max"])
def test_meter_matches_reference_testmeter(method):
    import sys
    import types

    ref_shim.install()
    for name in ("av", "matplotlib", "matplotlib.pyplot"):       # imported by the reference's utils, unused by the meter
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    from slowfast.utils.meters import TestMeter as RefMeter
    from slowfast.utils import metrics as ref_metrics

    nv, nc, ncls = 9, 4, 11
    batches = _batches(nv, nc, ncls, 5, seed=8)
    a = esf.TestMeter(nv, nc, ncls, len(batches), ensemble_method=method)
    b = RefMeter(nv, nc, ncls, len(batches), ensemble_method=method)
    for bt in batches:
        a.update_stats(*bt)
        b.update_stats(*bt)
    assert torch.allclose(a.video_preds, b.video_preds, atol=1e-6)
    assert torch.equal(a.video_labels, b.video_labels) and torch.equal(a.clip_count, b.clip_count)
    # the reference's own topks_correct (utils/metrics.py:40) calls .view(-1) on a non-contiguous slice and raises on
    # torch >= 2: its k = 1 branch still works and pins ours
    mine = esf.topks_correct(a.video_preds, a.video_labels, (1,))
    theirs = ref_metrics.topks_correct(b.video_preds, b.video_labels, (1,))
    assert [float(x) for x in mine] == [float(x) for x in theirs]


def test_checkpoint_roundtrip_and_key_normalisation(tmp_path):
    cfg = esf.slowfast_4x16_r50_cfg()
    cfg.NUM_GPUS = 0
    torch.manual_seed(1)
    src = esf.build_model(cfg)
    g = torch.Generator().manual_seed(5)
    for m in src.modules():                      # non-trivial BN statistics so a dropped key is noticed
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.num_batches_tracked.fill_(3)
    torch.manual_seed(2)
    dst = esf.build_model(cfg)
    path = str(tmp_path / "checkpoint_epoch_00007.pyth")
    esf.save_checkpoint(path, src, epoch=7)
    # the reference's own call shape (tools/test_net.py -> cu.load_test_checkpoint): positional data_parallel, optimizer
    assert esf.load_checkpoint(path, dst, False, None, inflation=False, convert_from_caffe2=False) == 7
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    # DDP prefix + Sub-BN naming as written by a multi-GPU training run of the reference
    sd = {}
    for k, v in src.state_dict().items():
        head, leaf = k.rsplit(".", 1)
        is_bn = leaf in ("running_mean", "running_var", "num_batches_tracked")
        if not is_bn:
            sd["module." + k] = v
        elif leaf == "num_batches_tracked":
            sd["module.%s.split_bn.%s" % (head, leaf)] = v
        else:
            sd["module.%s.bn.%s" % (head, leaf)] = v
            sd["module.%s.split_bn.%s" % (head, leaf)] = v.repeat(2)
    torch.save({"epoch": 3, "model_state": sd, "optimizer_state": {}, "cfg": ""}, path)
    torch.manual_seed(4)
    dst2 = esf.build_model(cfg)
    assert esf.load_checkpoint(path, dst2, data_parallel=False) == 3
    # reference defaults: data_parallel=True expects the DDP wrapper (model.module), inflation is out of scope
    import pytest
    with pytest.raises(AttributeError):
        esf.load_checkpoint(path, dst2)
    with pytest.raises(NotImplementedError):
        esf.load_checkpoint(path, dst2, False, None, inflation=True)
    wrapped = torch.nn.Module()
    wrapped.module = dst2
    assert esf.load_checkpoint(path, wrapped) == 3
    assert esf.load_checkpoint.last_report == {"missing": [], "unexpected": []}
    for (k, a), (_, b) in zip(src.state_dict().items(), dst2.state_dict().items()):
        assert torch.equal(a, b), k
