"""instruction mix and stall reasons of one kernel from `ncu -i rep --page source --csv` output (dev helper)"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2 and r[0].startswith("0x")]
op = collections.Counter(); stall = collections.Counter(); samp = collections.Counter()
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for r in data:
    src = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    o = m.group(2) if m else "?"
    o = ".".join(o.split(".")[:2]) if o.startswith(("LD", "ST", "ATOM", "RED")) else o.split(".")[0]
    op[o] += int(r[idx["Instructions Executed"]] or 0)
    samp[o] += int(r[idx["# Samples"]] or 0)
    for c in stall_cols:
        stall[c] += int(r[idx[c]] or 0)
T = sum(op.values()); S = sum(stall.values())
print("static SASS lines", len(data), " executed warp instructions", T)
for o, n in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{o:12s} {n:11d} {100 * n / T:5.1f}%   stall samples {samp[o]}")
for c, n in stall.most_common(10):
    print(f"{c:28s} {n:8d} {100 * n / S:5.1f}%")
