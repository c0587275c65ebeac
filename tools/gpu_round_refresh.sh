#!/bin/bash
# One GPU call that regenerates the round's evidence: parity suite, headline bench + reference arm, launch list + full
# ncu capture of the dominant kernel (tools/gpu_profile_round.sh), all five BASELINE configs, launch lists of configs 3/5.
# usage: tools/gpu_round_refresh.sh <tag>
TAG=${1:-r01x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
bash tools/gpu_profile_round.sh ${TAG} | tail -1
python tools/bench_configs.py --steps 10 --json gpurun_out/${TAG}_configs.json 2>&1 | tail -9
for k in 3 5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 34 -c 17 --csv --log-file gpurun_out/${TAG}_cfg${k}_launches.csv python tools/run_config_once.py $k > /dev/null 2>&1
done
python tools/bench_widening.py 2>&1 | tail -5 | tee gpurun_out/${TAG}_widening.jsonl
