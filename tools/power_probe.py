#!/usr/bin/env python
"""Board power and SM clock while (a) the graph-replayed train step, (b) one fused-projection GEMM and (c) one LayerNorm
kernel run back to back for a few seconds each: is the step bounded by the power cap?  """
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cara_b200 import train as T, kernels as K

SECS = float(os.environ.get("SECS", "3"))
dev = torch.device("cuda", 0)


def phase(name, fn, unit_ms=None):
    fn(); torch.cuda.synchronize()
    s = bench.ClockSampler(0); s.start()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < SECS:
        for _ in range(8):
            fn()
        n += 8
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    c = s.stop()
    print("%-34s %8.3f ms/iter  sm %s / %s MHz  power %s / %s W  reasons %s" % (
        name, e0.elapsed_time(e1) / n, c["sm_mhz"], c["sm_max_mhz"], c.get("power_w"), c.get("power_limit_w"), c["reasons"]))


bench._install_init_module()
cfg = bench.CONFIGS["vitb16_r16"]
vit, opt = bench.build_model(cfg, dev)
x = torch.randn(256, 3, 224, 224, device=dev); y = torch.randint(0, 100, (256,), device=dev)
step = T.GraphedStep(vit, opt, x, y)
phase("train step (graph replay)", lambda: step(x, y))
M = 50432
a = torch.randn(M, 3072, device=dev).bfloat16(); w = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16()
phase("GEMM fc2 N768 K3072", lambda: K.gemm_cp(a, w))
a2 = torch.randn(M, 768, device=dev).bfloat16(); w2 = (torch.randn(768, 768, device=dev) * 0.03).bfloat16()
phase("GEMM proj N768 K768", lambda: K.gemm_cp(a2, w2))
xr = torch.randn(M, 768, device=dev); dl = torch.randn(M, 768, device=dev).bfloat16()
g = torch.ones(768, device=dev); b = torch.zeros(768, device=dev)
phase("ln_fwd (HBM-bound)", lambda: K.ln_fwd(xr, g, b, delta=dl))
qkv = torch.randn(256, 197, 3, 12, 64, device=dev).bfloat16()
phase("attn_fwd", lambda: K.attn_fwd(qkv.view(-1), 256, 197, 12, 64, 0.125))
