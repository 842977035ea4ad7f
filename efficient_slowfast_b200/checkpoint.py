"""Checkpoint wire format of the reference: `.pyth` files written by
`slowfast/utils/checkpoint.py:122-160` are `torch.save({"epoch", "model_state", "optimizer_state", "cfg"})`; the model
state uses the same keys as this package's modules (tests/test_oracle_vs_reference.py), optionally behind a DDP
`module.` prefix, and Sub-BN checkpoints carry `bn.bn.*` / `bn.split_bn.*` pairs that collapse to plain `bn.*`
(checkpoint.py:290-324).  Caffe2 `.pkl` conversion and 2D->3D inflation are not on the eval path and not built."""
import torch


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


def load_checkpoint(path_to_checkpoint, model, data_parallel=False, strict=False):
    """Loads `model_state` into `model` (or `model.module` when `data_parallel`); returns the stored epoch or -1.
    Changing the weights invalidates the model's compiled launch plans (they fold BN into the packed weights)."""
    ckpt = torch.load(path_to_checkpoint, map_location="cpu", weights_only=False)
    state = ckpt["model_state"] if isinstance(ckpt, dict) and "model_state" in ckpt else ckpt
    target = model.module if data_parallel else model
    missing, unexpected = target.load_state_dict(_normalise_keys(state), strict=strict)
    if hasattr(target, "invalidate_plans"):
        target.invalidate_plans()
    load_checkpoint.last_report = {"missing": list(missing), "unexpected": list(unexpected)}
    return int(ckpt.get("epoch", -1)) if isinstance(ckpt, dict) else -1


def save_checkpoint(path, model, epoch=-1, cfg=None):
    """Same container the reference writes (utils/checkpoint.py:122-160), model state only."""
    sd = model.module.state_dict() if hasattr(model, "module") else model.state_dict()
    torch.save({"epoch": epoch, "model_state": {k: v.detach().cpu() for k, v in sd.items()},
                "cfg": cfg.dump() if hasattr(cfg, "dump") else None}, path)
