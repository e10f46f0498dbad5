#!/usr/bin/env python
"""Run the attention core fwd + bwd at the bench shape a few times (for ncu) and time them (L2 flushed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
B, N, H, D = 256, int(os.environ.get("N", 197)), 12, 64
qkv = torch.randn(B, N, 3, H, D, device="cuda").to(torch.bfloat16)
d_o = torch.randn(B * N, H * D, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, D ** -0.5)
    g = K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, D ** -0.5)
torch.cuda.synchronize()
if os.environ.get("ATTN_TIME", "1") == "1":
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    def t(fn, n=10):
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / n * 1e3
    print("attn_fwd %.1f us   attn_bwd (incl. delta pre-pass) %.1f us" % (
        t(lambda: K.attn_fwd(qkv.view(-1), B, N, H, D, D ** -0.5)),
        t(lambda: K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, D ** -0.5))))
print("ok")
