timeout 500 python -m pytest tests -m gpu -x -q -k "parity_case or properties or sharded or culling or headline or config or graph_replay or idempotent or moving" 2>&1 | tail -3
timeout 300 bash tools/gpu_cfgs.sh r03d 2,3,5
