"""Counts of the Blackwell-specific SASS mnemonics per kernel of librlg_b200.so (cuobjdump -sass): proof that the kernels
issue tcgen05 MMAs (UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMA tensor loads (UTMALDG), bulk copies (UBLKCP).
    python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan-rl_3d_b200", "lib", "librlg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDUX", "FFMA2", "FMNMX3", "HMMA",
      "REDG.E.ADD.F32x4", "REDG.E.ADD.F32x2", "REDG.E.ADD.64"]      # (vector float / 64-bit integer reductions of the Chamfer backward)
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.x]+)", line)
    if m:
        op = m.group(1)
        total[cur] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)} (sm_100a): instruction counts per kernel; only kernels using a listed mnemonic")
for fn, c in counts.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name[:80]:80s} total {total[fn]:6d}  " + "  ".join(f"{k} {v}" for k, v in c.items()))
