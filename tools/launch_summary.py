#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, average and share."""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows:
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki].replace("(anonymous namespace)::", "")))
    name = name.replace("at::native::", "native::")[:72]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print("%d launches, %.2f ms of kernel time (cold-cache, serialised: compare SHARES)" % (len(rows), T / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-74s launches %5d  avg %8.1f us  share %5.1f%%" % (k, cnt[k], v / cnt[k], 100 * v / T))
