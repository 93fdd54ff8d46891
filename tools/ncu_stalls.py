"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` dump.
    ncu -i rep.ncu-rep --page source --csv -k regex:<kernel> > src.csv ; python tools/ncu_stalls.py src.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
idx = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break                      # only the first kernel instance of the dump
    data.append(r)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("instructions", len(data), "samples", tot)


def op(r):
    s = re.sub(r"^@!?U?P\d+\s+", "", r[idx["Source"]].strip())
    return s.split()[0].split(".")[0]


byop, bystall, executed = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    byop[op(r)] += int(r[idx["# Samples"]])
    executed[op(r)] += int(r[idx["Instructions Executed"]])
    for s in stalls:
        bystall[s] += int(r[idx[s]])
print("samples by opcode:", [(k, v, executed[k]) for k, v in byop.most_common(16)])
print("samples by reason:", bystall.most_common(10))
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
    print(data.index(r), r[idx["Source"]].strip()[:90], "| samples", r[idx["# Samples"]],
          {s[6:]: r[idx[s]] for s in stalls if int(r[idx[s]]) > 4}, "| exec", r[idx["Instructions Executed"]])
