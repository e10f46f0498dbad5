#!/usr/bin/env python
"""Run one fused-projection GEMM shape a few times (for ncu)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M, N, K0 = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (50432, 3072, 768))]
pair = int(sys.argv[4]) if len(sys.argv) > 4 else 1
a = torch.randn(M, K0, device="cuda").bfloat16(); w = (torch.randn(N, K0, device="cuda") * 0.03).bfloat16()
b = torch.randn(N, device="cuda")
for _ in range(3):
    y = K.gemm_cp(a, w, bias=b, pair=pair)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    y = K.gemm_cp(a, w, bias=b, pair=pair)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("M=%d N=%d K=%d pair=%d: %.3f ms %.1f TFLOP/s" % (M, N, K0, pair, ms, 2.0 * M * N * K0 / ms / 1e9))
