#!/usr/bin/env python
"""Stall samples and executed warp instructions of one captured kernel, aggregated per CUDA source line (SASS offsets of the
ncu source page joined with nvdisasm -g line info of the library).  usage: tools/ncu_by_line.py rep K lib.so mangled-substring [N]"""
import csv, glob, os, re, subprocess, sys, tempfile
rep, K, lib, sub = sys.argv[1], int(sys.argv[2]), os.path.abspath(sys.argv[3]), sys.argv[4]
N = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[K]; hi = starts[K + 1] if K + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr = rows[lo + 1]; body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(body[0][ia], 16)
per_off = {int(r[ia], 16) - base: (int(r[isamp]), int(r[iex])) for r in body}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
line_of = {}
for cubin in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    m = re.search(r"\.section\s+\.text\.(\S*" + re.escape(sub) + r"\S*),", txt)
    if not m:
        continue
    sec = txt[m.start():]
    nxt = sec.find(".section", 10)
    sec = sec[:nxt] if nxt > 0 else sec
    cur = None
    for l in sec.splitlines():
        f = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.search(r"/\*([0-9a-f]{4,})\*/", l)
        if a and cur:
            line_of[int(a.group(1), 16)] = cur
    break
agg = {}
for off, (s, ex) in per_off.items():
    k = line_of.get(off, ("?", 0))
    a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += ex
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
src = {}
def text(k):
    if k[0] == "?": return ""
    if k[0] not in src:
        p = [q for q in glob.glob(os.path.dirname(lib) + "/" + k[0])]
        src[k[0]] = open(p[0]).read().splitlines() if p else []
    L = src[k[0]]
    return L[k[1] - 1].strip()[:110] if 0 < k[1] <= len(L) else ""
print(f"samples {ts}, warp instructions {te/1e6:.1f} M; top lines by samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{100*v[0]/ts:5.1f}% smp {100*v[1]/te:5.1f}% ins  {k[0]}:{k[1]:<5d} {text(k)}")
