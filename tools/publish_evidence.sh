#!/bin/bash
# Copies one gpu_round_refresh capture (gpurun_out/<tag>_*) into the tracked profiles/<round>_* files.
# usage: tools/publish_evidence.sh <tag> [round]
set -e
cd "$(dirname "$0")/.."
TAG=$1; RND=${2:-r02}
cp gpurun_out/${TAG}_configs.json profiles/${RND}_configs.json
for k in 3 5; do
  cp gpurun_out/${TAG}_cfg${k}_kernels.md profiles/${RND}_cfg${k}_kernels_ncu.md
  cp gpurun_out/${TAG}_cfg${k}_launches.csv profiles/${RND}_cfg${k}_launches.csv
done
cp gpurun_out/${TAG}_cfg35_launchlist.txt profiles/${RND}_cfg35_launchlist.txt
cp gpurun_out/${TAG}_widening.jsonl profiles/${RND}_widening_ops.jsonl
cp gpurun_out/${TAG}_pytest.txt profiles/${RND}_pytest_gpu.txt
python tools/make_profile_summary.py ${TAG} ${RND}
