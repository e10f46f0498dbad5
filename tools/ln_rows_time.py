#!/usr/bin/env python
"""LayerNorm + rank-R row contraction: the fused kernels (csrc/ln_rows.cu) against the stand-alone pair
(ln_fwd + adapter_rows_fwd, ln_bwd + adapter_rows_bwd) at the bench shape, L2 flushed before every launch group."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, C, N, Rp = 50432, int(os.environ.get("C", 768)), 197, 16
S = int(os.environ.get("S", 3))
ONCE = os.environ.get("LN_ONCE") == "1"      # ncu mode: one launch per configuration
x = torch.randn(M, C, device="cuda"); delta = torch.randn(M, C, device="cuda").bfloat16()
gamma = torch.randn(C, device="cuda"); beta = torch.randn(C, device="cuda")
rs = torch.ones(M // N, device="cuda")
dh = torch.randn(M, C, device="cuda").bfloat16(); dx_in = torch.randn(M, C, device="cuda")
a_t2 = (torch.randn(2 * Rp, C, device="cuda") * 0.1).bfloat16()
cs = torch.randn(S, Rp, device="cuda"); cs1 = cs[:1].contiguous()
T = torch.randn(M, Rp, device="cuda"); dc = torch.zeros(1, Rp, device="cuda")
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
xo, h, mean, rstd = K.ln_fwd(x, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N)

def t(fn, n=10):
    if ONCE:
        fn(); torch.cuda.synchronize(); return 0.0
    fn(); fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

def fwd_pair():
    _, hh, _, _ = K.ln_fwd(x, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N)
    K.adapter_rows_fwd(hh, a_t2, cs)
def bwd_pair():
    _, g = K.ln_bwd(dh, xo, mean, rstd, gamma, dx_in=dx_in, rowscale=rs, rows_per_sample=N, want_g=True)
    K.adapter_rows_bwd(g, a_t2, cs1, T, dsc=dc)
print("ln_fwd alone            %6.1f us" % t(lambda: K.ln_fwd(x, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N)))
print("ln_fwd + rows_fwd       %6.1f us" % t(fwd_pair))
print("ln_fwd_rows (fused)     %6.1f us" % t(lambda: K.ln_fwd_rows(x, gamma, beta, a_t2, cs, delta=delta, rowscale=rs, rows_per_sample=N)))
print("ln_bwd alone            %6.1f us" % t(lambda: K.ln_bwd(dh, xo, mean, rstd, gamma, dx_in=dx_in, rowscale=rs, rows_per_sample=N, want_g=True)))
print("ln_bwd + rows_bwd       %6.1f us" % t(bwd_pair))
print("ln_bwd_rows (fused)     %6.1f us" % t(lambda: K.ln_bwd_rows(dh, xo, mean, rstd, gamma, a_t2, cs1, T, dc, dx_in=dx_in, rowscale=rs, rows_per_sample=N)))
