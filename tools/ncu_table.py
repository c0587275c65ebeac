#!/usr/bin/env python
"""One row per captured kernel of an ncu raw CSV (ncu -i x.ncu-rep --page raw --csv): duration, DRAM traffic, issue / occupancy /
L1TEX utilisation and the top stall reasons.  usage: tools/ncu_table.py gpurun_out/x_raw.csv [--md]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
md = "--md" in sys.argv
def scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
out = []
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    def f(k):
        try: return float(d[k][0].replace(",", "")) * scale(d[k][1])
        except Exception: return float("nan")
    st = sorted(((f(h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in d
                 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h), reverse=True)[:3]
    name = d["Kernel Name"][0]
    name = name[:name.index("(")] if "(" in name else name
    out.append((name, f("gpu__time_duration.sum"), f("dram__bytes_read.sum") / 1e6, f("dram__bytes_write.sum") / 1e6, f("smsp__inst_executed.sum") / 1e6,
                f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("sm__warps_active.avg.pct_of_peak_sustained_active"),
                f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), d.get("launch__registers_per_thread", ("?",))[0],
                d.get("launch__grid_size", ("?",))[0], " ".join(f"{n}={v:.2f}" for v, n in st)))
if md:
    print("| kernel | us | DRAM read MB | DRAM write MB | warp instr M | issue active % | warps active % | L1TEX LSU pipe % | regs | grid | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for o in out:
        print(f"| {o[0]} | {o[1]:.1f} | {o[2]:.1f} | {o[3]:.1f} | {o[4]:.2f} | {o[5]:.1f} | {o[6]:.1f} | {o[7]:.1f} | {o[8]} | {o[9]} | {o[10]} |")
else:
    for o in out:
        print(f"{o[0][:44]:44s} {o[1]:8.1f}us rd {o[2]:8.1f} wr {o[3]:8.1f} MB inst {o[4]:8.2f}M issue {o[5]:5.1f}% warps {o[6]:5.1f}% l1pipe {o[7]:5.1f}% regs {o[8]:>3s} grid {o[9]:>6s} | {o[10]}")
