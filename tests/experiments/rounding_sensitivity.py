"""CPU experiment (test infrastructure, not collected by pytest): how much does 16-bit STORAGE rounding alone move the
output of a golden case?  Runs the oracle twice -- exact FP32, and with activations rounded to fp16 / bf16 after every
ReLU ("act"), inside the Non-local blocks ("nl") or both -- and prints max|d| / max|ref| of the probabilities.

    python tests/experiments/rounding_sensitivity.py i3d_nln_r50 s96 act [float16|bfloat16]

Measured (FP16): i3d_nln_r50 act 1.5e-2, nl 2.2e-3, both 1.6e-2; i3d_r50 act 3.2e-4; slow_nln_r50 both 2.6e-3.
"""
import sys, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for q in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, q)
import helpers, recipe
from oracle import slowfast_oracle as O
import torch.nn.functional as F
name, tag = sys.argv[1], sys.argv[2]
mode = sys.argv[3]   # none | act | nl | both
dt = torch.float16 if len(sys.argv) < 5 else getattr(torch, sys.argv[4])
cfg, model, gold = helpers.case_model_and_weights(name)
xs = helpers.case_inputs(name, tag)
sd = model.state_dict()
y0 = O.forward(cfg, sd, xs)
rnd = lambda t: t.to(dt).to(t.dtype)
orig_relu = F.relu
orig_nl = O.nonlocal_block
if mode in ("act","both"):
    O.F.relu = lambda x, *a, **k: rnd(orig_relu(x))
if mode in ("nl","both"):
    oc = O._conv
    def nl(x, sd_, p, pool, inst):
        def conv(x_, s_, q, **kw):
            y = oc(x_, s_, q, **kw)
            return rnd(y) if any(t in q for t in ("conv_theta","conv_phi","conv_g")) else y
        O._conv = conv
        osm = F.softmax
        O.F.softmax = lambda t, dim=None: rnd(osm(t, dim=dim))
        oe = torch.einsum
        def es(eq, ops):
            r = oe(eq, ops)
            return rnd(r) if eq.startswith("ntg") else r
        torch.einsum = es
        try:
            return rnd(orig_nl(x, sd_, p, pool, inst)) if False else orig_nl(x, sd_, p, pool, inst)
        finally:
            O._conv = oc; O.F.softmax = osm; torch.einsum = oe
    O.nonlocal_block = nl
y1 = O.forward(cfg, sd, xs)
print(name, mode, dt, "rel err %.3e" % helpers.rel_err(y1, y0), " vs golden %.3e" % helpers.rel_err(y0, gold[tag+"/probs"]))
