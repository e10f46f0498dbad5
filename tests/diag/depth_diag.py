#!/usr/bin/env python
"""Diagnostics (test infrastructure, not collected by pytest): bf16 path vs the fp32 SIMT mode on the GPU as a function of
depth / width / rank.  python tests/diag/depth_diag.py"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cara_oracle as O  # noqa: E402
from tests.test_parity_gpu import build, run_step, rel, cos  # noqa: E402

warnings.simplefilter("ignore")


def case(geom, batch=4, scale=1.0):
    from cara_b200.fp32 import set_precision
    g = O.Geometry(**geom)
    x, y = O.synthetic_batch(g, batch)
    vit, _ = build(g, scale)
    vit.eval()
    logits, loss, grads = run_step(vit, x, y)
    del vit
    ref, _ = build(g, scale)
    set_precision(ref, "fp32")
    ref.eval()
    rl, rloss, rg = run_step(ref, x, y)
    del ref
    torch.cuda.empty_cache()
    worst = min((cos(grads[k], rg[k]), k) for k in rg)
    print("%-70s logits rel %.3e  |logit| %.3f  worst grad cos %.6f (%s)" %
          (str(geom), rel(logits, rl), float(rl.abs().mean()), worst[0], worst[1]), flush=True)


if __name__ == "__main__":
    for depth in (2, 6, 12, 24):
        case(dict(embed_dim=1024, depth=depth, num_heads=16, rank=32, num_classes=100))
    case(dict(embed_dim=1024, depth=24, num_heads=16, rank=16, num_classes=100))
    case(dict(embed_dim=1024, depth=24, num_heads=16, rank=32, num_classes=100), scale=0.0)
    case(dict(embed_dim=768, depth=24, num_heads=12, rank=16, num_classes=100))
    case(dict(embed_dim=768, depth=12, num_heads=12, rank=32, num_classes=100))
    case(dict(embed_dim=1280, depth=32, num_heads=16, patch=14, rank=32, num_classes=100), batch=3)
