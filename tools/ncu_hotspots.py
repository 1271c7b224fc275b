"""Stall-sample summary of an `ncu --page source --csv` export (gzip or plain): totals per stall reason, executed
warp-instructions per opcode, and the instructions that collect the most samples.
usage: python tools/ncu_hotspots.py gpurun_out/<kernel>_source.csv.gz [blocks]     (blocks: divide executed counts by it)"""
import collections
import csv
import gzip
import io
import sys

path = sys.argv[1]
blocks = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
fh = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(fh))
print("kernel:", rows[0][1][:140])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
c = collections.Counter()
for r in data:
    for s in stalls:
        c[s] += int(r[ix[s]])
print(f"SASS instructions {len(data)}, warp samples {tot}")
print("stall reasons:", ", ".join(f"{k[6:]} {100.0 * v / tot:.1f}%" for k, v in c.most_common(9)))


def op(s):
    t = s.strip().split()
    return t[1] if t[0].startswith("@") else t[0]


oc, sc = collections.Counter(), collections.Counter()
for r in data:
    oc[op(r[ix["Source"]])] += int(r[ix["Instructions Executed"]])
    sc[op(r[ix["Source"]])] += int(r[ix["# Samples"]])
print(f"executed warp-instructions per block ({blocks:g} blocks): {sum(oc.values()) / blocks:.1f}")
print("  " + ", ".join(f"{o} {n / blocks:.1f}" for o, n in oc.most_common(18)))
print("hottest instructions (index, SASS, samples, executed, top stall reasons):")
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:25]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {i:5d} {r[ix['Source']].strip()[:60]:60s} {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>9s}  " + ", ".join(f"{n} {k}" for n, k in st))
