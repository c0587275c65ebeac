#!/bin/bash
cd vk_tessellated_clusters_b200/csrc
for v in "$@"; do
  make clean > /dev/null; make EXTRA="$v" > /dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  (cd ../..; echo "== $v"; python tools/bench_configs.py --steps 8 --only 2,5,3 --no-oracle 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:12], round(d['ms'],4), {k: round(x,3) for k,x in d['stage_ms'].items() if x>0.03})")
done
make clean > /dev/null; make > /dev/null 2>&1
