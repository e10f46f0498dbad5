#!/usr/bin/env python
"""Run the small kernels around the hot path once each at the bench shapes (for ncu): eval-merge reconstruction
(ViT-B fc1: 3072 x 768, rank 16, 4 slices), fused AdamW over the 122 k-element trainable buffer (device-state variant),
patch im2col and token assembly at batch 256, the CP-factor operand staging."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
dev = "cuda"
W = torch.randn(3072, 768, device=dev) * 0.02
A = torch.randn(768, 16, device=dev) * 0.1; Bf = torch.randn(768, 16, device=dev) * 0.1; cs = torch.randn(4, 16, device=dev)
n = 121924
p, g, m, v = (torch.randn(n, device=dev) for _ in range(4)); v = v.abs()
state = torch.tensor([1e-3, 0.0, 0.0, 0.0], device=dev)
img = torch.randn(256, 3, 224, 224, device=dev)
cls = torch.randn(768, device=dev); pos = torch.randn(197, 768, device=dev)
F = torch.randn(12, 3072, 16, device=dev)
for _ in range(2):
    w = K.merge_weights(W, A, Bf, cs)
    K.adamw_step_dev(p, g, m, v, state)
    pt = K.patchify(img, 16, 768)
    x = K.assemble_tokens(pt, cls, pos, 256, 197, 768)
    K.factor_operands(F, 16)
torch.cuda.synchronize()
print("ok")
