#!/bin/bash
# per-kernel durations (+ instructions, DRAM bytes) of one frame of the given configs.  usage: tools/gpu_launchlist.sh <tag> "3 5"
TAG=$1
mkdir -p gpurun_out
for k in $2; do
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -c 64 --csv --log-file gpurun_out/${TAG}_cfg${k}_launches.csv python tools/run_config_once.py $k > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${TAG}_cfg${k}_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[hdr]; iN=H.index("Kernel Name"); iM=H.index("Metric Name"); iV=H.index("Metric Value"); iI=H.index("ID")
d={}
for r in rows[hdr+1:]:
    if len(r)<=iV: continue
    d.setdefault((int(r[iI]), r[iN].split("(")[0][:40]),{})[r[iM]]=float(r[iV].replace(",",""))
print("== config $k")
tot=0
items=sorted(d.items())
last=max(j for j,((i,n),m) in enumerate(items) if "k_frame_begin" in n)
for (i,n),m in items[last:]:
    t=m.get("gpu__time_duration.sum",0)/1e3; tot+=t
    print(f"{n:42s} {t:8.1f} us  inst {m.get('smsp__inst_executed.sum',0)/1e6:8.2f} M  rd {m.get('dram__bytes_read.sum',0)/1e6:8.1f} MB  wr {m.get('dram__bytes_write.sum',0)/1e6:8.1f} MB")
print("total", round(tot,1))
PY
done
