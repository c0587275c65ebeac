#!/bin/bash
# `ncu --set full` captures (with source) of the frame's kernels on the headline workload and on configs 3 / 5.
# usage: tools/gpu_capture.sh <tag> [what: 2,3,5]      -> gpurun_out/<tag>_cfgK.ncu-rep (+ _raw.csv)
TAG=${1:-r02}
WHAT=${2:-2,3,5}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
if [[ ",$WHAT," == *",2,"* ]]; then
  # headline: 2 warm-up frames x 11 matching launches are skipped, one frame is captured
  $NCU -k regex:"k_instantiate|k_triangle_split|k_cluster_classify|k_blas" -s 22 -c 11 -o gpurun_out/${TAG}_cfg2 -f python tools/run_config_once.py 2 > /dev/null 2>&1
fi
for k in 3 5; do
  if [[ ",$WHAT," == *",$k,"* ]]; then
    $NCU -k regex:"k_cluster_classify|k_classify_scan|k_class_cache|k_cluster_copies|k_cluster_vertices|k_mini_vertices|k_instantiate" -s 14 -c 7 -o gpurun_out/${TAG}_cfg$k -f python tools/run_config_once.py $k > /dev/null 2>&1
  fi
done
for r in gpurun_out/${TAG}_cfg*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
done
ls -la gpurun_out/${TAG}_cfg*
