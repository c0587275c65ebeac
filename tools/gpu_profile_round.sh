#!/bin/bash
# Produces the per-round evidence under gpurun_out/ (copied into profiles/ afterwards): bench line, per-launch ncu
# durations of the same command, one full ncu capture of the dominant kernel.  Usage: tools/gpu_profile_round.sh r01
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_instantiate -s 3 -c 1 -o gpurun_out/${TAG}_instantiate python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_instantiate.ncu-rep --page details > gpurun_out/${TAG}_instantiate_details.txt 2>/dev/null
ncu -i gpurun_out/${TAG}_instantiate.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]; vals=rows[2]
keep=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for h,u,v in zip(hdr,units,vals):
    if h in keep: print(f'{h},{u},{v}')
" > gpurun_out/${TAG}_instantiate_raw.csv
cat gpurun_out/${TAG}_bench.json
