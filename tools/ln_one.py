#!/usr/bin/env python
"""Run the fused residual-add + LayerNorm forward / backward at the bench shape (for ncu / timing)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, C, N = 50432, 768, 197
x = torch.randn(M, C, device="cuda"); delta = torch.randn(M, C, device="cuda").bfloat16()
gamma = torch.randn(C, device="cuda"); beta = torch.randn(C, device="cuda")
rs = torch.ones(M // N, device="cuda")
dh = torch.randn(M, C, device="cuda").bfloat16(); dx_in = torch.randn(M, C, device="cuda")
def run():
    xo, h, mean, rstd = K.ln_fwd(x, gamma, beta, delta=delta, rowscale=rs, rows_per_sample=N)
    K.ln_bwd(dh, xo, mean, rstd, gamma, dx_in=dx_in, rowscale=rs, rows_per_sample=N, want_g=True)
for _ in range(2):
    run()
torch.cuda.synchronize()
e = [torch.cuda.Event(True) for _ in range(2)]
e[0].record()
for _ in range(5):
    run()
e[1].record(); torch.cuda.synchronize()
print("ln fwd+bwd pair: %.1f us" % (e[0].elapsed_time(e[1]) * 200))
