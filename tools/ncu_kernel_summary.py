"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV): python tools/ncu_kernel_summary.py file.csv [steps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, cnt = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].replace("void ", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([\w:]+)", name)
    n = m.group(1).split("::")[-1] if m else name[:40]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    agg[n] += v
    cnt[n] += 1
tot = sum(agg.values())
print(f"total {tot / steps / 1e3:.2f} ms per step, {sum(cnt.values()) // steps} launches per step")
for n, v in agg.most_common(45):
    print(f"{n:40s} {cnt[n] // steps:5d} {v / steps / 1e3:8.3f} ms {v / tot:6.3f}")
