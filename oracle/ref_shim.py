"""TEST INFRASTRUCTURE ONLY -- container-side import shim for the upstream reference.

Imports weidafeng/Efficient-SlowFast's own PyTorch model zoo from /root/reference
(read-only mount, exists only in the build container, never on the GPU box) so
that (1) `oracle/slowfast_oracle.py` can be validated against the real reference
forward and (2) `tests/golden/make_golden.py` can generate golden vectors.

The reference imports a few packages that are not installed in this image
(yacs via fvcore.common.config, detectron2, mmcv, simplejson, portalocker); none
of them is used by the classification forward path, so they are replaced with
minimal stand-ins (recipe: SURVEY.md Appendix B).

Nothing in the product package imports this module.
"""
import ast
import contextlib
import copy
import io
import json
import os
import sys
import types

import torch.nn as nn
import yaml

REF_ROOT = os.environ.get("ESF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "SlowFast", "slowfast"))


class CfgNode(dict):
    """attr-dict stand-in for yacs/fvcore CfgNode (only what defaults.py needs)."""

    def __init__(self, d=None, **_):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k]._merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, f):
        with open(f) as fh:
            self._merge(yaml.safe_load(fh))

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node, ks = self, k.split(".")
            for kk in ks[:-1]:
                node = node[kk]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except Exception:
                    pass
            node[ks[-1]] = v


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.path.insert(0, os.path.join(REF_ROOT, "SlowFast"))
    sys.path.insert(0, os.path.join(REF_ROOT, "config_slowfast", "fvcore"))

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    _mod("fvcore.common.config", CfgNode=CfgNode)

    class _ROIAlign:
        def __init__(self, *a, **k):
            raise NotImplementedError("detection head is out of scope")

    _mod("detectron2", layers=_mod("detectron2.layers", ROIAlign=_ROIAlign))

    def _constant_init(m, val, bias=0):
        if getattr(m, "weight", None) is not None:
            nn.init.constant_(m.weight, val)
        if getattr(m, "bias", None) is not None:
            nn.init.constant_(m.bias, bias)

    def _kaiming_init(m, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
        nn.init.kaiming_normal_(m.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if getattr(m, "bias", None) is not None:
            nn.init.constant_(m.bias, bias)

    _mod("mmcv", cnn=_mod("mmcv.cnn", constant_init=_constant_init, kaiming_init=_kaiming_init))
    _mod("simplejson", dumps=lambda o, sort_keys=False, use_decimal=False, **k: json.dumps(o, sort_keys=sort_keys, default=str))
    _mod("portalocker")
    _installed = True


def get_cfg(yaml_rel=None, opts=()):
    """Reference default cfg (+ optional YAML under SlowFast/configs, + KEY VALUE opts)."""
    install()
    from slowfast.config.defaults import get_cfg as _get

    cfg = _get()
    if yaml_rel:
        cfg.merge_from_file(os.path.join(REF_ROOT, "SlowFast", yaml_rel))
    if opts:
        cfg.merge_from_list(list(opts))
    cfg.NUM_GPUS = 0
    return cfg


def build_reference_model(cfg):
    """The reference's own build_model(cfg) (CPU), constructor prints silenced."""
    install()
    import slowfast.models.build as B
    import slowfast.models  # noqa: F401  (registers the classes)

    with contextlib.redirect_stdout(io.StringIO()):
        model = B.build_model(cfg)
    return model


def reference_dataset_utils():
    """slowfast/datasets/utils.py of the reference (tensor_normalize, pack_pathway_output) without executing the
    package __init__, which imports the video decoders (`av`) that this image does not have."""
    install()
    import slowfast  # noqa: F401
    name = "slowfast.datasets"
    if name not in sys.modules or not hasattr(sys.modules[name], "__path__"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_ROOT, "SlowFast", "slowfast", "datasets")]
        sys.modules[name] = pkg
    import importlib
    return importlib.import_module("slowfast.datasets.utils")
