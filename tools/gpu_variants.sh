#!/bin/bash
# rebuild the library with different compile-time knobs on the GPU box and time the headline bench for each
# usage: tools/gpu_variants.sh "<EXTRA flags 1>" "<EXTRA flags 2>" ...
cd vk_tessellated_clusters_b200/csrc
for v in "$@"; do
  make clean > /dev/null; make EXTRA="$v" > /dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  (cd ../..; python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '| frame', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['stage_ms'].items()})")
done
make clean > /dev/null; make > /dev/null 2>&1
