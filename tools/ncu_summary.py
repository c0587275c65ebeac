#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: duration, issue/pipe utilisation, stall reasons, memory wavefronts.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [row]"""
import csv, subprocess, sys
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def f(k):
    try: return float(d[k][0].replace(",", ""))
    except Exception: return float("nan")
print("kernel", d.get("Kernel Name", ("?",))[0], "grid", d.get("launch__grid_size", ("?",))[0], "regs", d.get("launch__registers_per_thread", ("?",))[0])
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum"]:
    if k in d: print(f"  {k:90s} {d[k][0]:>16s} {d[k][1]}")
print("pipes (pct of peak, active):")
for h in sorted(d):
    if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active") and f(h) > 0.5:
        print(f"  {h[len('sm__inst_executed_pipe_'):-len('.avg.pct_of_peak_sustained_active')]:24s} {f(h):6.1f}")
print("stalls (warps per issue):")
st = [(f(h), h) for h in d if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for v, h in sorted(st, reverse=True):
    if v > 0.05: print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
