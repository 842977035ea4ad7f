"""Oracle vs the real reference forward, stage by stage (build container only: needs /root/reference)."""
import pytest
import torch

import helpers
import recipe
from oracle import ref_shim
from oracle import slowfast_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("name", ["dual_r50", "slowfast_r50", "shufflenetv2_w05", "shufflenet_w2g3", "mobilenetv2_w1",
                                  "ghostnet_w1", "i3d_r50", "slow_r50", "slow_nln_r50", "i3d_nln_r50", "slowfast_r50_fcn"])
def test_every_stage_bit_exact(name):
    spec = recipe.CASES[name]
    cfg = ref_shim.get_cfg(spec["yaml"], spec["opts"])
    torch.manual_seed(0)
    ref = ref_shim.build_reference_model(cfg)
    gold = helpers.load_golden(name)
    bn = {k[3:]: v for k, v in gold.items() if k.startswith("bn/")}
    ref.load_state_dict(recipe.seeded_state_dict(ref.state_dict(), seed=0, bn_stats=bn,
                                                  stress=spec.get("stress", False)), strict=True)
    ref.eval()
    if spec.get("single"):      # fixed head pool: the clip must have the cfg's frames / crop
        xs = [recipe.seeded_clip(1, cfg.DATA.NUM_FRAMES, cfg.DATA.CROP_SIZE, seed=5)]
    else:
        xs = recipe.pack_pathway_output(recipe.seeded_clip(2, 32, 48, seed=5), cfg.SLOWFAST.ALPHA)
    got = {}
    hooks = [getattr(ref, n).register_forward_hook(lambda m, i, o, n=n: got.__setitem__(n, [t.clone() for t in o]))
             for n in ("s0", "s1", "s1_fuse", "s2", "s2_fuse", "s3", "s3_fuse", "s4", "s4_fuse", "s5", "s5_fuse", "s6",
                       "s7", "s7_fuse", "s8") if hasattr(ref, n)]
    with torch.no_grad():
        y_ref = ref([t.clone() for t in xs])
    for h in hooks:
        h.remove()
    taps = {}
    y = O.forward(cfg, ref.state_dict(), xs, taps=taps)
    for n, ts in got.items():
        for pw in range(len(ts)):
            assert torch.equal(ts[pw], taps[n][pw]), (n, pw)
    assert torch.equal(y, y_ref)


@pytest.mark.parametrize("name", ["dual_r50", "slowfast_r50", "shufflenetv2_w05", "shufflenet_w2g3", "mobilenetv2_w1",
                                  "ghostnet_w1", "i3d_r50", "slow_r50", "slow_nln_r50", "i3d_nln_r50", "slowfast_r50_fcn"])
def test_state_dict_schema_and_seeded_init_match_reference(name):
    import efficient_slowfast_b200 as esf

    spec = recipe.CASES[name]
    rcfg = ref_shim.get_cfg(spec["yaml"], spec["opts"])
    torch.manual_seed(11)
    ref = ref_shim.build_reference_model(rcfg)
    torch.manual_seed(11)
    mine = esf.build_model(helpers.case_cfg(name))
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in b:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    assert [n for n, _ in mine.named_children()] == [n for n, _ in ref.named_children()]
    # a reference checkpoint loads into the drop-in (utils/checkpoint.py:279 uses strict=False)
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_caffe2_name_fixture_is_the_reference_mapping():
    """tests/golden/caffe2_names.json was produced by the reference's own converter (utils/c2_model_loading.py)."""
    import importlib.util
    import json
    import os

    ref_shim.install()
    spec = importlib.util.spec_from_file_location(
        "c2_model_loading", os.path.join(ref_shim.REF_ROOT, "SlowFast", "slowfast", "utils", "c2_model_loading.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    convert = mod.get_name_convert_func()
    want = json.load(open(os.path.join(helpers.GOLDEN_DIR, "caffe2_names.json")))
    for c2, key in want.items():
        assert convert(c2) == key


def _config_signature(path):
    import yaml
    with open(path) as fh:
        y = yaml.safe_load(fh)
    m, r, d = y.get("MODEL", {}), y.get("RESNET", {}), y.get("DATA", {})
    return (m.get("MODEL_NAME"), m.get("ARCH"), r.get("DEPTH"), r.get("WIDTH_PER_GROUP"), str(d.get("INPUT_CHANNEL_NUM")),
            str(y.get("NONLOCAL", {}).get("LOCATION")), y.get("SLOWFAST", {}).get("WIDTH_MULTI"))


def test_reference_yaml_configs_build_with_the_same_schema():
    """Every classification YAML the reference itself can build (57 of its 72: detection configs are out of scope, 8
    name unregistered models or fail its own asserts) must build here with the same state_dict keys and shapes and, for
    the same seed, the same initial weights.  One YAML per distinct (model, arch, depth, width, input channels,
    non-local layout) signature by default; ESF_FULL_CONFIG_SWEEP=1 runs all of them (6 min;
    profiles/r2_reference_config_sweep.txt is that run)."""
    import glob
    import os

    import efficient_slowfast_b200 as esf

    root = os.path.join(ref_shim.REF_ROOT, "SlowFast")
    seen, checked = set(), 0
    for y in sorted(glob.glob(os.path.join(root, "configs", "**", "*.yaml"), recursive=True)):
        sig = _config_signature(y)
        if sig in seen and not os.environ.get("ESF_FULL_CONFIG_SWEEP"):
            continue
        seen.add(sig)
        try:
            cfg = ref_shim.get_cfg(os.path.relpath(y, root), [])
            if cfg.DETECTION.ENABLE:
                continue
            torch.manual_seed(0)
            ref = ref_shim.build_reference_model(cfg)
        except Exception:
            continue                     # the reference cannot build this YAML itself
        ours_cfg = esf.get_cfg()
        ours_cfg.merge_from_file(y)
        ours_cfg.NUM_GPUS = 0
        torch.manual_seed(0)
        ours = esf.build_model(ours_cfg)
        a, b = ref.state_dict(), ours.state_dict()
        assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}, y
        assert all(torch.equal(a[k], b[k]) for k in a), y
        checked += 1
    assert checked >= 12
