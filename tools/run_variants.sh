#!/bin/bash
# times the headline bench (and optionally other configs) for every csrc/variants/libtess_*.so; usage: tools/run_variants.sh [configs e.g. 2,5]
cd "$(dirname "$0")/.."
for so in vk_tessellated_clusters_b200/csrc/libtess_clusters.so vk_tessellated_clusters_b200/csrc/variants/libtess_*.so; do
  n=$(basename $so .so)
  TC_LIB_PATH=$PWD/$so python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', '| frame', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['stage_ms'].items()})"
  if [ -n "$1" ]; then
    TC_LIB_PATH=$PWD/$so python tools/bench_configs.py --steps 8 --only $1 --no-oracle 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['config'][:12], round(d['ms'],4), {k: round(x,3) for k,x in d['stage_ms'].items() if x>0.03})"
  fi
done
