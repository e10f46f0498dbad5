#!/usr/bin/env python
"""Diagnostics (test infrastructure, not collected by pytest): where does the bf16 path's error against the fp32
oracle come from?  python tests/diag/parity_diag.py"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cara_oracle as O  # noqa: E402
from tests.test_parity_gpu import build, run_step, rel, cos  # noqa: E402

warnings.simplefilter("ignore")


def case(tag, geom, batch, scale, drop_path=0.0, train=False):
    g = O.Geometry(**geom)
    vit, st = build(g, scale, drop_path=drop_path)
    vit.train(train)
    keep = None
    if train and drop_path > 0:
        from cara_b200 import vit as V
        drawn = []
        orig = V.DropPath.rowscale

        def spy(self, b, device):
            rs = orig(self, b, device)
            drawn.append(None if rs is None else rs.detach().cpu())
            return rs
        V.DropPath.rowscale = spy
    x, y = O.synthetic_batch(g, batch)
    logits, loss, grads = run_step(vit, x, y)
    if train and drop_path > 0:
        V.DropPath.rowscale = orig
        keep = torch.ones(g.depth, 2, batch)
        it = iter(drawn)
        for l in range(g.depth):
            for j in range(2):
                if isinstance(vit.blocks[l].drop_path, V.DropPath):
                    rs = next(it)
                    if rs is not None:
                        keep[l, j] = rs
        print("   keep:", keep.flatten().tolist())
    ol, oloss, og = O.loss_and_grads(st, g, x, y, scale, keep=keep)
    print("%s: logits rel %.3e  loss %.5f vs %.5f" % (tag, rel(logits, ol), loss, float(oloss)))
    print("   per-sample rel:", ["%.2e" % rel(logits[i], ol[i]) for i in range(batch)])
    print("   grad cos:", {k: "%.5f" % cos(grads[k], og[k]) for k in og})
    print("   grad rel:", {k: "%.1e" % rel(grads[k], og[k]) for k in og})


if __name__ == "__main__":
    case("d2 r8 s2.5", dict(depth=2, rank=8, num_classes=10), 3, 2.5)
    case("d2 r8 s1.0", dict(depth=2, rank=8, num_classes=10), 3, 1.0)
    case("d3 r16 nodp", dict(depth=3, rank=16, num_classes=10), 4, 1.0)
    case("d3 r16 dp.5", dict(depth=3, rank=16, num_classes=10), 4, 1.0, drop_path=0.5, train=True)
    case("d12 r16", dict(depth=12, rank=16, num_classes=100), 4, 1.0)
