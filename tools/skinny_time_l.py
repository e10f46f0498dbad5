#!/usr/bin/env python
"""Skinny adapter kernels at the ViT-L/16 rank-32 shapes (M = 50432, C = 1024, Rp = 32), L2 flushed between launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, Rp, C = 50432, 32, 1024
flush = torch.empty(512 << 20, device="cuda", dtype=torch.uint8)
def t(fn, n=6):
    fn(); fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3
bf = lambda *s: torch.randn(*s, device="cuda").bfloat16()
for Kd, S in [(C, 3), (C, 1), (C, 4), (4 * C, 1)]:
    x = bf(M, Kd); a_t2 = bf(2 * Rp, Kd); sc = torch.randn(S, Rp, device="cuda")
    us = t(lambda: K.adapter_rows_fwd(x, a_t2, sc))
    print("rows_fwd  K=%4d S=%d: %6.1f us  %5.2f TB/s" % (Kd, S, us, (M * Kd * 2 + M * Rp * (4 + 6 * S)) / us / 1e6))
for N, S in [(3 * C, 3), (C, 1), (4 * C, 4)]:
    g = bf(M, N); b_t2 = bf(2 * Rp, N // S); sc = torch.randn(S, Rp, device="cuda"); T = torch.randn(M, Rp, device="cuda")
    us = t(lambda: K.adapter_rows_bwd(g, b_t2, sc, T))
    print("rows_bwd  N=%4d S=%d: %6.1f us  %5.2f TB/s" % (N, S, us, (M * N * 2 + M * Rp * 10) / us / 1e6))
for Kc, S, cs in [(C, 1, False), (4 * C, 1, False), (3 * C, 3, False), (4 * C, 4, True)]:
    x = bf(M, Kc); v = bf(M, S * 3 * Rp)
    out = torch.zeros(Kc // S, Rp, device="cuda"); csum = torch.zeros(Kc, device="cuda") if cs else None
    us = t(lambda: K.adapter_cols(x, v, S, Rp, want_colsum=cs, out=out, cs=csum))
    print("cols      K=%4d S=%d cs=%d: %6.1f us  %5.2f TB/s" % (Kc, S, cs, us, (M * Kc * 2 + M * S * 3 * Rp * 2) / us / 1e6))
