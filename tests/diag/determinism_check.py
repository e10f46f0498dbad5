#!/usr/bin/env python
"""How reproducible is one backward?  (test infrastructure, not collected by pytest: python tests/diag/determinism_check.py)

Four identical eager steps of fresh models in one process; the flat gradients must agree to ~1e-7 (the order of the
fp32 atomics in the factor-gradient kernels).  History: an atomically reduced split-K in the head's small GEMMs put
1e-7 noise into the LOGITS, which flipped bf16 roundings all the way down the backward pass and showed up as 2e-4
differences between runs; the split-K partial sums are now reduced in a fixed order (csrc/misc.cu)."""
import os, sys, warnings, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
warnings.simplefilter("ignore")
from tests.test_parity_gpu import build, rel
from oracle import cara_oracle as O   # test helper only (synthetic state / batch)
from cara_b200 import train as T

g = O.Geometry(depth=2, rank=8, num_classes=10)
x, y = O.synthetic_batch(g, 4, seed=100)
x, y = x.cuda(), y.cuda()
def grads(graphed):
    vit, _ = build(g, 1.0); vit.train()
    opt = T.FusedAdamW(T.FlatTrainable(T.freeze_backbone(vit)), lr=1e-3)
    if graphed:
        step = T.GraphedStep(vit, opt, x, y); step(x, y)
    else:
        T.train_step(vit, opt, x, y)
    return opt.flat.grad.detach().cpu().clone(), opt.flat.slices
runs = [("eager0", False), ("eager1", False), ("eager2", False), ("eager3", False)]
G = {}
for name, gr in runs:
    G[name], sl = grads(gr)
names = [n for n, _ in runs]
print("pairwise rel differences of the flat gradient:")
for i, a in enumerate(names):
    print("  %-7s" % a + " ".join("%9.2e" % rel(G[a], G[b]) for b in names[:i]))
