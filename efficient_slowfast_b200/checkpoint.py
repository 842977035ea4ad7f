"""Checkpoint wire format of the reference: `.pyth` files written by
`slowfast/utils/checkpoint.py:122-160` are `torch.save({"epoch", "model_state", "optimizer_state", "cfg"})`; the model
state uses the same keys as this package's modules (tests/test_oracle_vs_reference.py), optionally behind a DDP
`module.` prefix, and Sub-BN checkpoints carry `bn.bn.*` / `bn.split_bn.*` pairs that collapse to plain `bn.*`
(checkpoint.py:290-324).  Caffe2 `.pkl` checkpoints (the model zoo's format: `{"blobs": {caffe2 name: ndarray}}`,
checkpoint.py:206-259) are converted by name (`caffe2_to_pytorch_name`, same mapping as utils/c2_model_loading.py).
2D->3D inflation is a training-time feature and not built."""
import os
import pickle
import re

import numpy as np
import torch

# Caffe2 blob name -> PyTorch key (utils/c2_model_loading.py:10-112 defines the mapping as a chain of substitutions; it
# is restated here as a structural parse: pathway prefix, block address, role inside the block, parameter leaf).
_LEAF = {"w": "weight", "b": "bias"}
_BN_LEAF = {"s": "weight", "w": "weight", "b": "bias", "rm": "running_mean", "riv": "running_var"}
_NL_ROLE = {"theta": "conv_theta", "phi": "conv_phi", "g": "conv_g", "out": "conv_out"}
_RE_NONLOCAL = re.compile(r"^nonlocal_conv(\d+)_(\d+)_(theta|phi|g|out|bn)_(.+)$")
_RE_FUSE = re.compile(r"^t_(?:pool1|res(\d+)_\d+_branch2c_bn)_subsample_(bn_)?(.+)$")
_RE_BLOCK = re.compile(r"^(t_)?res(\d+)_(\d+)_branch(\d+)([a-z]?)_(.+)$")
_RE_STEM = re.compile(r"^(t_)?(?:res_conv1|conv1)_(.+)$")


def _leaf(rest, bn):
    """'w' / 'b' / 'bn_s' ... plus Caffe2's optimizer suffixes (e.g. 'w_momentum'), which stay unmapped."""
    table = _BN_LEAF if bn else _LEAF
    return table.get(rest, rest)


def caffe2_to_pytorch_name(name):
    """Caffe2 blob name of the SlowFast model zoo -> state_dict key of the model classes in this package."""
    m = _RE_NONLOCAL.match(name)
    if m:
        stage, idx, role, rest = m.groups()
        base = "s%s.pathway0_nonlocal%s." % (stage, idx)
        return base + ("bn." + _leaf(rest, True) if role == "bn" else _NL_ROLE[role] + "." + _leaf(rest, False))
    m = _RE_FUSE.match(name)
    if m:
        stage, bn, rest = m.groups()
        base = "s%s_fuse." % (stage or "1")
        return base + ("bn." + _leaf(rest, True) if bn else "conv_f2s." + _leaf(rest, False))
    m = _RE_BLOCK.match(name)
    if m:
        fast, stage, idx, branch, letter, rest = m.groups()
        base = "s%s.pathway%d_res%s.branch%s" % (stage, 1 if fast else 0, idx, branch)
        sep = "." + letter + "_" if letter else "_"            # branch2.a_bn.weight  vs  branch1_bn.weight
        if rest.startswith("bn_"):
            return base + sep + "bn." + _leaf(rest[3:], True)
        tail = _leaf(rest, False)
        return base + ("." + letter if letter else "") + "." + tail if tail != rest else base + sep + rest
    m = _RE_STEM.match(name)
    if m:
        fast, rest = m.groups()
        base = "s1.pathway%d_stem." % (1 if fast else 0)
        if rest.startswith("bn_"):
            return base + "bn." + _leaf(rest[3:], True)
        return base + "conv." + _leaf(rest, False)
    if name.startswith("pred_"):
        return "head.projection." + _leaf(name[5:], False)
    return name


def convert_caffe2_blobs(blobs, model_state):
    """checkpoint.py:211-258: converted blobs whose shape matches the model's tensor (1-D blobs that divide a longer
    1-D model tensor are tiled: normal-BN statistics into a Sub-BN model).  Returns (state_dict, skipped names)."""
    out, skipped = {}, []
    for key, blob in blobs.items():
        name = caffe2_to_pytorch_name(key)
        if name not in model_state:
            if not any(t in key for t in ("momentum", "lr", "model_iter")):
                skipped.append(key)
            continue
        blob = np.asarray(blob)
        want = tuple(model_state[name].shape)
        if blob.ndim == 1 and len(want) == 1 and want[0] > blob.shape[0] and want[0] % blob.shape[0] == 0:
            blob = np.concatenate([blob] * (want[0] // blob.shape[0]))
        if tuple(blob.shape) == want:
            out[name] = torch.tensor(blob).clone()
        else:
            skipped.append(key)
    return out, skipped


def _normalise_keys(state):
    """DDP prefix off; Sub-BN layout -> plain BN (the inverse of what training writes, checkpoint.py:290-324):
    `X.bn.running_{mean,var}` -> `X.running_{mean,var}`, `X.split_bn.num_batches_tracked` -> `X.num_batches_tracked`,
    the other `split_bn` entries are training-time copies and dropped; affine weight / bias already live on `X`."""
    sub_bn = any(".split_bn." in k for k in state)     # only Sub-BN checkpoints carry the extra `.bn` level
    out = {}
    for k, v in state.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if ".split_bn." in k:
            if not k.endswith(".split_bn.num_batches_tracked"):
                continue
            k = k.replace(".split_bn.", ".")
        elif sub_bn and k.endswith((".bn.running_mean", ".bn.running_var")):
            head, _, leaf = k.rsplit(".", 2)
            k = head + "." + leaf
        out[k] = v
    return out


def load_checkpoint(path_to_checkpoint, model, data_parallel=True, optimizer=None, inflation=False,
                    convert_from_caffe2=False, *, strict=False):
    """utils/checkpoint.py:178-285 of the reference, same positional order, names and defaults (its call sites pass
    `load_checkpoint(path, model, cfg.NUM_GPUS > 1, None, inflation=False, convert_from_caffe2=...)`).  Loads
    `model_state` into `model.module` when `data_parallel` (the DDP wrapper, as in the reference: a bare model with
    data_parallel=True raises AttributeError there too) else into `model`; restores `optimizer` from
    `optimizer_state` when given; returns the stored epoch or -1 (always -1 for a Caffe2 file).  `inflation`
    (2D -> 3D weight inflation, a fine-tuning feature) is out of scope and raises.  `strict` is keyword-only and not
    part of the reference signature (the reference always loads with strict=False).  Changing the weights invalidates
    the model's compiled launch plans (they fold BN into the packed weights)."""
    if not os.path.exists(path_to_checkpoint):
        raise AssertionError("Checkpoint '{}' not found".format(path_to_checkpoint))
    if inflation:
        raise NotImplementedError("inflation of 2D checkpoints (utils/checkpoint.py:271-276) is a training-time feature "
                                  "outside the forward path")
    target = model.module if data_parallel else model
    if convert_from_caffe2:
        with open(path_to_checkpoint, "rb") as f:
            c2 = pickle.load(f, encoding="latin1")
        state, skipped = convert_caffe2_blobs(c2["blobs"], target.state_dict())
        missing, unexpected = target.load_state_dict(state, strict=False)
        if hasattr(target, "invalidate_plans"):
            target.invalidate_plans()
        load_checkpoint.last_report = {"missing": list(missing), "unexpected": list(unexpected), "skipped": skipped}
        return -1
    ckpt = torch.load(path_to_checkpoint, map_location="cpu", weights_only=False)
    state = ckpt["model_state"] if isinstance(ckpt, dict) and "model_state" in ckpt else ckpt
    missing, unexpected = target.load_state_dict(_normalise_keys(state), strict=strict)
    if hasattr(target, "invalidate_plans"):
        target.invalidate_plans()
    if optimizer:
        optimizer.load_state_dict(ckpt["optimizer_state"])
    load_checkpoint.last_report = {"missing": list(missing), "unexpected": list(unexpected)}
    return int(ckpt.get("epoch", -1)) if isinstance(ckpt, dict) else -1


def save_checkpoint(path, model, epoch=-1, cfg=None):
    """Same container the reference writes (utils/checkpoint.py:122-160), model state only."""
    sd = model.module.state_dict() if hasattr(model, "module") else model.state_dict()
    torch.save({"epoch": epoch, "model_state": {k: v.detach().cpu() for k, v in sd.items()},
                "cfg": cfg.dump() if hasattr(cfg, "dump") else None}, path)
