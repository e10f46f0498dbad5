#!/usr/bin/env python
"""Run one fused-projection GEMM shape a few times (for ncu / timing): gemm_one.py M N K epi [ext]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K, _lib as L
M, N, K0 = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (50432, 3072, 768))]
epi = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ext = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a = torch.randn(M, K0, device="cuda").bfloat16(); w = (torch.randn(N, K0, device="cuda") * 0.03).bfloat16()
b = torch.randn(N, device="cuda")
kw = {}
if ext:
    S = 4 if N == 3072 else (3 if N == 2304 else 1)
    kw = dict(a1=torch.randn(M, S * 48, device="cuda").bfloat16(), b1=torch.randn(N // S, 48, device="cuda").bfloat16(), ext_slices=S)
if epi == 2:
    kw["aux"] = torch.randn(M, N, device="cuda").bfloat16()
if epi == 3:   # EPI_DELTA: the output projection's dX GEMM also writes the attention backward's row term
    kw["delta"] = (torch.randn(M, N, device="cuda").bfloat16(), (torch.randn(M, N, device="cuda") * 0.01).bfloat16(),
                   torch.empty(M // 197, N // 64, 197, device="cuda"), 197)
    b = None
run = lambda: K.gemm_cp(a, w, bias=b, epi=epi, **kw)  # noqa: E731
for _ in range(3):
    y = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    y = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("M=%d N=%d K=%d epi=%d ext=%d: %.3f ms %.1f TFLOP/s" % (M, N, K0, epi, ext, ms, 2.0 * M * N * (K0 + (16 if ext else 0)) / ms / 1e9))
