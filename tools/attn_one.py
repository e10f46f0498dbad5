#!/usr/bin/env python
"""Run the attention core fwd + bwd at the bench shape a few times (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
B, N, H, D = 256, 197, 12, 64
qkv = torch.randn(B, N, 3, H, D, device="cuda").to(torch.bfloat16)
d_o = torch.randn(B * N, H * D, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, D ** -0.5)
    g = K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, D ** -0.5)
torch.cuda.synchronize()
print("ok")
