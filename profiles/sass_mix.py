"""Aggregate an `ncu --page source --csv` export by SASS opcode: executed warp instructions per warp and stall samples.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python profiles/sass_mix.py src.csv WARPS"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = heads[0]
if len(heads) > 1:          # the export repeats the listing per view (SASS, then source-correlated): keep the first
    rows = rows[:heads[1]]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops, samples = collections.Counter(), collections.Counter()
tot = totsamp = 0
stall_names = [h for h in hdr if h.startswith("stall_")]
stalls = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    sass = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    base = (m.group(2) if m else sass).split(".")[0]
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    ops[base] += n
    samples[base] += s
    tot += n
    totsamp += s
    for h in stall_names:
        v = r[ix[h]]
        if v.isdigit():
            stalls[h] += int(v)
print(f"total warp instructions {tot}  per warp {tot / warps:.1f}   static SASS instructions {len(rows) - hi - 1}")
for k, v in ops.most_common(22):
    print(f"{k:12s} {v / warps:9.1f} /warp {100 * v / tot:5.1f}%   stall samples {100 * samples[k] / max(totsamp, 1):5.1f}%")
print("stall reasons (all samples):")
for k, v in stalls.most_common(8):
    print(f"  {k:40s} {100 * v / max(sum(stalls.values()), 1):5.1f}%")
