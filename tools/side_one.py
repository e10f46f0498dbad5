#!/usr/bin/env python
"""One stand-alone side-tile launch (for ncu): T = X A, Uhat at the bench shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, Kd, S, Rp = 50432, int(os.environ.get("KD", 768)), int(os.environ.get("S", 1)), 16
x = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
P = (torch.randn(2 * Rp, Kd, device="cuda") * 0.1).to(torch.bfloat16)
sc = torch.randn(S, Rp, device="cuda")
for _ in range(3):
    T, U = K.adapter_rows_fwd(x, P, sc)
torch.cuda.synchronize()
