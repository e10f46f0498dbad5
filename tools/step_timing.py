#!/usr/bin/env python
"""Host enqueue time vs device time of one train step (is the step launch-bound?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cara_b200 import train as T, kernels as K

bench._install_init_module()
cfg = bench.CONFIGS["vitb16_r16"]
dev = torch.device("cuda", 0)
vit, opt = bench.build_model(cfg, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 3, 224, 224, device=dev); y = torch.randint(0, 100, (B,), device=dev)
for _ in range(3):
    T.train_step(vit, opt, x, y)
torch.cuda.synchronize()
for it in range(3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    out = vit(x); t1 = time.perf_counter()
    loss = torch.nn.functional.cross_entropy(out, y)
    opt.zero_grad(); loss.backward(); t2 = time.perf_counter()
    opt.step(); e1.record(); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print("B=%d host: fwd %.1f ms, bwd %.1f ms, opt %.2f ms, total enqueue %.1f ms; wait-for-GPU %.1f ms; device span %.1f ms"
          % (B, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t3-t0)*1e3, (t4-t3)*1e3, e0.elapsed_time(e1)))
