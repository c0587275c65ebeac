#!/bin/bash
# parity suite + headline bench line on one GPU.  usage: tools/gpu_check.sh <tag> [pytest args]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
