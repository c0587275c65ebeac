#!/usr/bin/env python
"""Top stalled SASS instructions of a capture. usage: tools/ncu_hot_sass.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the page holds one block per captured kernel ("Kernel Name" row, header row, body); take block K (argv[3], default 0)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[K]; hi = starts[K + 1] if K + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr = rows[lo + 1]; body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in body)
base = int(body[0][ia], 16)
recs = []
for k, r in enumerate(body):
    s = int(r[isamp]); 
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    recs.append((s, k, int(r[ia], 16) - base, r[isrc].strip(), r[iex], st))
print("total samples", tot)
for s, k, off, src, ex, st in sorted(recs, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  {off:#06x} ex={ex:>9s} {src[:70]:70s} {st}")
# stall by region of 64 instructions
print("by 32-instruction block:")
for b in range(0, len(recs), 32):
    s = sum(r[0] for r in recs[b:b+32])
    if s / tot > 0.01: print(f"  {recs[b][2]:#06x}: {100*s/tot:5.1f}%")
