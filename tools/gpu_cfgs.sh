#!/bin/bash
# quick parity subset + configs timing for the default library and the variants.  usage: tools/gpu_cfgs.sh <tag> <configs e.g. 2,3,5> [pytest -k expr]
TAG=$1; CFGS=$2
mkdir -p gpurun_out
if [ -n "$3" ]; then timeout 1200 python -m pytest tests -m gpu -x -q -k "$3" 2>&1 | tail -6; fi
for so in vk_tessellated_clusters_b200/csrc/libtess_clusters.so vk_tessellated_clusters_b200/csrc/variants/libtess_*.so; do
  [ -f $so ] || continue
  echo "== $(basename $so .so)"
  TC_LIB_PATH=$PWD/$so timeout 600 python tools/bench_configs.py --steps 10 --only $CFGS 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:12], 'ms', round(d['ms'],4), 'frac', round(d['frac_measured'],3), {k: round(x,3) for k,x in d['stage_ms'].items()})" | tee -a gpurun_out/${TAG}_cfgs.txt
done
