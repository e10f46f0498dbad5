#!/usr/bin/env python
"""Per-kernel device time inside a real train step, without ncu: CUDA events around every C-ABI launch
(warm caches, real clocks, kernels still back to back).  Prints avg time per (entry point, shape) and its share."""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cara_b200 import train as T, kernels as K, _lib as L

bench._install_init_module()
cfg = bench.CONFIGS[os.environ.get("CFG", "vitb16_r16")]
dev = torch.device("cuda", 0)
vit, opt = bench.build_model(cfg, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["batch"]
x = torch.randn(B, 3, 224, 224, device=dev); y = torch.randint(0, 100, (B,), device=dev)
for _ in range(3):
    T.train_step(vit, opt, x, y)
torch.cuda.synchronize()

trace, pending, tag = [], [None], [""]
orig_prep, orig_check, orig_gemm = K._prep, L.check, K.gemm_cp
def prep(t):
    st = orig_prep(t)
    e = torch.cuda.Event(enable_timing=True); e.record(); pending[0] = e
    return st
def check(rc, what):
    orig_check(rc, what)
    e = torch.cuda.Event(enable_timing=True); e.record()
    trace.append((what + tag[0], pending[0], e)); tag[0] = ""
def gemm(a0, b0, *args, **kw):
    tag[0] = " M%d N%d K%d epi%d%s" % (a0.shape[0], b0.shape[0], a0.shape[1], kw.get("epi", 0), (" +ext" if kw.get("a1") is not None else "") + (" +side" if kw.get("side") is not None else ""))
    return orig_gemm(a0, b0, *args, **kw)
K._prep, L.check, K.gemm_cp = prep, check, gemm
steps = 3
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(steps):
    T.train_step(vit, opt, x, y)
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / steps
agg = collections.OrderedDict()
for name, a, b in trace:
    v = agg.setdefault(name, [0, 0.0]); v[0] += 1; v[1] += a.elapsed_time(b)
ours = sum(v[1] for v in agg.values()) / steps
print("step %.2f ms (with event overhead); inside cara kernels %.2f ms; %d launches/step" % (total, ours, len(trace) // steps))
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s n/step %4d  avg %8.1f us  per-step %6.2f ms  %5.1f%%" % (name, v[0] // steps, 1e3 * v[1] / v[0], v[1] / steps, 100 * v[1] / steps / total))
