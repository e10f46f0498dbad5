#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) into one text table: tools/ncu_summary.py rep1 rep2 ... > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__cycles_elapsed.max", "cycles")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print("%s: unreadable" % rep)
        continue
    hdr, units = rows[0], rows[1]
    print("== %s" % rep.split("/")[-1])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        parts = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(r[i])
                    if label.endswith("_MB") or label == "time_us":
                        v *= SCALE.get(units[i], 1.0)
                    v = "%.1f" % v
                except ValueError:
                    v = r[i]
                parts.append("%s=%s" % (label, v))
        print("  %-40s %s" % (name[:40], " ".join(parts)))
