#!/usr/bin/env python
"""Time the skinny adapter kernels at the bench shapes (M = 50432)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, Rp = 50432, 16
ONCE = os.environ.get("SKINNY_ONCE") == "1"      # ncu mode: one launch per configuration
def t(fn, n=10):
    if ONCE:
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
bf = lambda *s: torch.randn(*s, device="cuda").bfloat16()
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for Kd, S in [(768, 3), (768, 1), (768, 4), (3072, 1)]:
    x = bf(M, Kd); a_t2 = bf(2 * Rp, Kd); sc = torch.randn(S, Rp, device="cuda")
    us = t(lambda: K.adapter_rows_fwd(x, a_t2, sc))
    print("rows_fwd  K=%4d S=%d: %6.1f us  %5.2f TB/s" % (Kd, S, us, (M * Kd * 2 + M * Rp * (4 + 6 * S)) / us / 1e6))
for N, S in [(2304, 3), (768, 1), (3072, 4), (768, 1)]:
    g = bf(M, N); b_t2 = bf(2 * Rp, N // S); sc = torch.randn(S, Rp, device="cuda"); T = torch.randn(M, Rp, device="cuda")
    us = t(lambda: K.adapter_rows_bwd(g, b_t2, sc, T))
    print("rows_bwd  N=%4d S=%d: %6.1f us  %5.2f TB/s" % (N, S, us, (M * N * 2 + M * Rp * 10) / us / 1e6))
for Kc, S, cs in [(768, 1, False), (3072, 1, False), (2304, 3, False), (3072, 4, True), (768, 1, True)]:
    x = bf(M, Kc); v = bf(M, S * 3 * Rp)
    us = t(lambda: K.adapter_cols(x, v, S, Rp, want_colsum=cs))
    print("cols      K=%4d S=%d cs=%d: %6.1f us  %5.2f TB/s" % (Kc, S, cs, us, (M * Kc * 2 + M * S * 3 * Rp * 2) / us / 1e6))
