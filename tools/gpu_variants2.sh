#!/bin/bash
# times the headline frame + stage times for the default library and every csrc/variants/libtess_*.so (tools/build_variants.sh)
# usage: tools/gpu_variants2.sh <tag> [pytest -k expression for a quick parity check of the default build]
TAG=${1:-var}
mkdir -p gpurun_out
if [ -n "$2" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 | tail -8; fi
for so in vk_tessellated_clusters_b200/csrc/libtess_clusters.so vk_tessellated_clusters_b200/csrc/variants/libtess_*.so; do
  n=$(basename $so .so)
  TC_LIB_PATH=$PWD/$so timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>gpurun_out/${TAG}_$n.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$n', '| frame median', round(d['frame_ms']['median'],4), 'mean', round(d['frame_ms']['mean'],4), {k: round(x,4) for k,x in d['stage_ms_median'].items()}, 'inst frac', round(d['roofline']['frac'],4))" 2>&1 | tee -a gpurun_out/${TAG}_variants.txt
done
