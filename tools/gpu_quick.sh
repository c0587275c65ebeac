#!/bin/bash
# quick GPU check: parity tests, headline bench (stage times), optional full ncu capture of k_instantiate.  usage: tools/gpu_quick.sh [tag]
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
if [ -n "$1" ]; then
  mkdir -p gpurun_out
  ncu --set full --clock-control none --import-source on -k regex:${2:-k_instantiate} -s ${3:-3} -c 1 -o gpurun_out/$1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
fi
