#!/usr/bin/env python
"""Per-kernel device time of the train step from torch.profiler (CUPTI), warm, in situ."""
import os, sys, collections, re
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cara_b200 import train as T

bench._install_init_module()
key = sys.argv[2] if len(sys.argv) > 2 else "vitb16_r16"
cfg = bench.CONFIGS[key]
dev = torch.device("cuda", 0)
vit, opt = bench.build_model(cfg, dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(B, 3, 224, 224, device=dev); y = torch.randint(0, 100, (B,), device=dev)
for _ in range(3):
    T.train_step(vit, opt, x, y)
torch.cuda.synchronize()
STEPS = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        T.train_step(vit, opt, x, y)
    torch.cuda.synchronize()
tot = collections.defaultdict(float); cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "")))[:64]
        tot[name] += ev.device_time; cnt[name] += 1
T_all = sum(tot.values())
print("device kernel time per step: %.2f ms (B=%d, %s)" % (T_all / STEPS / 1e3, B, key))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:22]:
    print("%-66s n/step %6.1f avg %8.1f us  per-step %7.2f ms %5.1f%%" % (k, cnt[k] / STEPS, v / cnt[k], v / STEPS / 1e3, 100 * v / T_all))
