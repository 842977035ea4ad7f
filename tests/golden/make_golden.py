"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_shim.py).  Build-container only; the fixtures it writes are committed and travel to the GPU box.

    python tests/golden/make_golden.py [case ...]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import recipe  # noqa: E402
from oracle import ref_shim  # noqa: E402


def calibrate_bn(model, clips, alpha):
    """precise-BN style: momentum=None (cumulative average) over train-mode forwards of the reference."""
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm3d)]
    for m in bns:
        m.reset_running_stats()
        m.momentum = None
    model.train()
    drops = [m for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    for d in drops:
        d.p_saved, d.p = d.p, 0.0
    with torch.no_grad():
        for x in clips:
            model([t.clone() for t in recipe.pack_pathway_output(x, alpha)])
    for d in drops:
        d.p = d.p_saved
    model.eval()


def patch_large_n_attention(threshold=40000, rows=2048):
    """SURVEY 8(c) "Large-N oracle": at N = T*H*W = 100 352 keys (SlowFastGhostNet, 224^2 x 32 frames) the reference's
    N x N affinity is 40 GB per clip and cannot be materialised.  SpatialAttention.forward
    (wdf_attention_helper.py:33-54) is replaced, for such inputs only, by the same arithmetic evaluated for `rows`
    query rows at a time: every row still sees its complete softmax over all N keys (bmm -> softmax(dim=-1) -> bmm),
    only the batching of the rows changes."""
    from slowfast.models import wdf_attention_helper as W

    orig = W.SpatialAttention.forward

    def forward(self, x):
        B, C, T, H, Wd = x.size()
        N = T * H * Wd
        if N <= threshold:
            return orig(self, x)
        q = self.query_conv(x).view(B, -1, N).permute(0, 2, 1)
        k = self.key_conv(x).view(B, -1, N)
        v = self.value_conv(x).view(B, -1, N)
        out = torch.empty(B, C, N, dtype=x.dtype)
        for r0 in range(0, N, rows):
            att = self.softmax(torch.bmm(q[:, r0:r0 + rows], k))
            out[:, :, r0:r0 + rows] = torch.bmm(v, att.permute(0, 2, 1))
        return self.gamma * out.view(B, C, T, H, Wd) + x

    W.SpatialAttention.forward = forward


def make_case(name):
    spec = recipe.CASES[name]
    ref_shim.install()
    patch_large_n_attention()
    cfg = ref_shim.get_cfg(spec["yaml"], spec["opts"])
    assert cfg.MODEL.MODEL_NAME == spec["model"]
    alpha = 0 if spec.get("single") else cfg.SLOWFAST.ALPHA
    torch.manual_seed(0)
    model = ref_shim.build_reference_model(cfg)
    sd = recipe.seeded_state_dict(model.state_dict(), seed=0, stress=spec.get("stress", False))
    model.load_state_dict(sd, strict=True)
    cb, cf, cs = spec["calib"]
    ch = spec.get("channels", 3)
    calibrate_bn(model, [recipe.seeded_clip(cb, cf, cs, seed=100 + i, channels=ch) for i in range(2)], alpha)
    out = {}
    for k, v in model.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["bn/" + k] = v.numpy().astype(np.float32)
    stages = [n for n, _ in model.named_children() if n.startswith("s") and n != "softmax"]
    for tag, b, frames, crop in spec["inputs"]:
        taps = {}
        hooks = []
        for sname in stages:
            hooks.append(getattr(model, sname).register_forward_hook(
                lambda m, i, o, sname=sname: taps.__setitem__(sname, [t.detach().clone() for t in o])))
        fc = [m for m in model.head.modules() if isinstance(m, torch.nn.Linear)][-1]
        hooks.append(fc.register_forward_hook(lambda m, i, o: taps.__setitem__("logits", o.detach().clone())))
        x = recipe.seeded_clip(b, frames, crop, seed=1, channels=ch)
        with torch.no_grad():
            y = model([t.clone() for t in recipe.pack_pathway_output(x, alpha)])
        for h in hooks:
            h.remove()
        out["%s/probs" % tag] = y.numpy().astype(np.float32)
        out["%s/logits" % tag] = taps["logits"].reshape(b, -1).numpy().astype(np.float32)
        for sname in stages:
            for pw, t in enumerate(taps[sname]):
                flat = t.reshape(-1)
                idx = recipe.sample_indices(flat.numel())
                out["%s/%s/%d/shape" % (tag, sname, pw)] = np.array(t.shape, dtype=np.int64)
                out["%s/%s/%d/stats" % (tag, sname, pw)] = np.array(
                    [flat.mean().item(), flat.std().item(), flat.abs().max().item()], dtype=np.float64)
                out["%s/%s/%d/samples" % (tag, sname, pw)] = flat[idx].numpy().astype(np.float32)
        top2 = torch.topk(y, 2, dim=1).values
        print(name, tag, "probs max %.4f  top1/top2 margin %.3f  |logit| max %.3f" % (
            y.max().item(), ((top2[:, 0] - top2[:, 1]) / top2[:, 0]).min().item(), taps["logits"].abs().max().item()))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def make_default_init():
    """Corner case: the reference's own seeded default init (gamma = 0, zero final BN, fresh BN stats)."""
    out = {}
    for name, spec in recipe.CASES.items():
        if spec.get("stress") or not name.endswith("_r50"):
            continue
        cfg = ref_shim.get_cfg(spec["yaml"], spec["opts"])
        torch.manual_seed(1234)
        model = ref_shim.build_reference_model(cfg).eval()
        x = recipe.seeded_clip(2, 32, 64, seed=1)
        with torch.no_grad():
            y = model([t.clone() for t in recipe.pack_pathway_output(x, cfg.SLOWFAST.ALPHA)])
        out[name + "/probs"] = y.numpy().astype(np.float32)
        print("default-init", name, y.max().item())
    np.savez_compressed(os.path.join(HERE, "default_init.npz"), **out)


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("reference tree not mounted; golden vectors can only be generated in the build container")
    names = sys.argv[1:] or list(recipe.CASES)
    for n in names:
        if n == "default_init":
            make_default_init()
        else:
            make_case(n)
    if not sys.argv[1:]:
        make_default_init()
