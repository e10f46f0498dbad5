#!/usr/bin/env python
"""Is the projection GEMM bound per SM (issue / ingest) or chip-wide (L2 -> SM fabric)?  Time a shape on 148 / 111 / 74 /
37 persistent CTAs: per-SM throughput that RISES as CTAs are removed means a shared (chip-level) limit."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cara_b200 import kernels as K
M = 50432
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
def t(fn, reps=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for name, N, K0 in (("fc2  N768 K3072", 768, 3072), ("proj N768 K768", 768, 768), ("qkv  N2304 K768", 2304, 768)):
    a = torch.randn(M, K0, device="cuda").bfloat16(); w = (torch.randn(N, K0, device="cuda") * 0.03).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for sms in (148, 111, 74, 37):
        us = t(lambda: K.gemm_cp(a, w, out=out, num_sms=sms))
        tf = 2.0 * M * N * K0 / us / 1e6
        print("%s  %3d CTAs: %7.1f us  %7.1f TFLOP/s  %.2f TFLOP/s per SM" % (name, sms, us, tf, tf / sms), flush=True)
