"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean device time.
    python tools/launch_summary.py gpurun_out/x_launches.csv [skip_first_n] [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
data = []
for r in rows[hdr + 1:]:
    if len(r) > vi:
        t = float(r[vi].replace(",", ""))
        t = t / 1e3 if r[ui] in ("ns", "nsecond") else (t * 1e3 if r[ui] in ("ms", "msecond") else t)     # -> us
        data.append((r[ki], t))
data = data[skip:]
tot = sum(t for _, t in data)
print(f"{len(data)} launches, {tot:.1f} us of kernel time (cold, serialised)")
agg = collections.OrderedDict()
for k, t in data:
    k = k.split("(")[0][:78]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f} %  x{c:4d}  mean {t / c:7.2f} us  {k}")
