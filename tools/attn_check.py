#!/usr/bin/env python
"""Attention core check + timing through the C ABI (CARA_ATTN_TC selects tcgen05 / mma.sync kernels)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
BF16 = torch.bfloat16
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())

def check(B, N, H, D, bwd=True):
    g = torch.Generator(device="cuda").manual_seed(6)
    C = H * D
    qkv = torch.randn(B, N, 3, H, D, device="cuda", generator=g).to(BF16)
    scale = D ** -0.5
    o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, scale)
    torch.cuda.synchronize()
    q, k, v = [t.float().requires_grad_(True) for t in qkv.permute(2, 0, 3, 1, 4)]
    att = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
    oref = (att @ v).transpose(1, 2).reshape(B * N, C)
    lref = torch.logsumexp((q @ k.transpose(-2, -1)) * scale, -1) / math.log(2.0)
    msg = "B%d N%d H%d D%d: o %.2e  o+lo %.2e  lse %.2e" % (B, N, H, D, rel(o.float(), oref.detach()),
          rel(o.float() + o_lo.float(), oref.detach()), rel(lse, lref.detach()))
    if bwd:
        d_o = torch.randn(B * N, C, device="cuda", generator=g).to(BF16)
        oref.backward(d_o.float())
        dqkv = K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, scale).view(B, N, 3, H, D)
        torch.cuda.synchronize()
        dref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4)
        msg += "  dq %.2e dk %.2e dv %.2e" % tuple(rel(dqkv[:, :, i].float(), dref[:, :, i]) for i in range(3))
    print(msg, flush=True)

def timeit(B, N, H, D, bwd=True):
    qkv = torch.randn(B, N, 3, H, D, device="cuda").to(BF16)
    d_o = torch.randn(B * N, H * D, device="cuda").to(BF16)
    scale = D ** -0.5
    for _ in range(3):
        o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, scale)
        if bwd:
            K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, scale)
    e = [torch.cuda.Event(True) for _ in range(3)]
    e[0].record()
    for _ in range(10):
        o, o_lo, lse = K.attn_fwd(qkv.view(-1), B, N, H, D, scale)
    e[1].record()
    if bwd:
        for _ in range(10):
            K.attn_bwd(qkv.view(-1), o, o_lo, lse, d_o, B, N, H, D, scale)
    e[2].record()
    torch.cuda.synchronize()
    print("B%d N%d H%d D%d: fwd %.1f us  bwd %.1f us" % (B, N, H, D, e[0].elapsed_time(e[1]) * 100, e[1].elapsed_time(e[2]) * 100), flush=True)

if __name__ == "__main__":
    bwd = "--nobwd" not in sys.argv
    for shp in [(1, 128, 1, 64), (1, 64, 2, 64), (2, 17, 3, 64), (1, 50, 2, 64), (3, 197, 12, 64), (2, 256, 2, 64), (1, 200, 3, 64)]:
        check(*shp, bwd=bwd)
    timeit(256, 197, 12, 64, bwd=bwd)
    import ctypes as C
    from cara_b200 import _lib as L
    buf = (C.c_longlong * 64)()
    if L.lib().cara_debug_read(buf, 64) and buf[0]:
        v = list(buf)
        base = v[0]
        print("stamps (cycles after stamp 0): %s" % [(i, v[i] - base) for i in range(1, 64) if v[i]])
