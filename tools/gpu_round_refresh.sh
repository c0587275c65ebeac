#!/bin/bash
# One GPU call that regenerates the round's evidence: parity suite, headline bench + reference arm, launch list + full
# ncu capture of the dominant kernel (tools/gpu_profile_round.sh), all five BASELINE configs, launch lists of configs 3/5.
# usage: tools/gpu_round_refresh.sh <tag>
TAG=${1:-r01x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
bash tools/gpu_profile_round.sh ${TAG} | tail -1
timeout 900 python tools/bench_configs.py --steps 10 --parity --json gpurun_out/${TAG}_configs.json 2>&1 | tail -9
timeout 300 bash tools/gpu_launchlist.sh ${TAG} "3 5" | tee gpurun_out/${TAG}_cfg35_launchlist.txt
timeout 600 bash tools/gpu_capture.sh ${TAG} 3,5 > /dev/null 2>&1
for k in 3 5; do python tools/ncu_table.py gpurun_out/${TAG}_cfg${k}_raw.csv --md > gpurun_out/${TAG}_cfg${k}_kernels.md; done
timeout 600 python tools/bench_widening.py 2>&1 | tail -5 | tee gpurun_out/${TAG}_widening.jsonl
