timeout 300 python -m pytest tests -m gpu -x -q -k "parity_case or config3 or config5 or sharded or culling or properties" 2>&1 | tail -3
timeout 300 bash tools/gpu_cfgs.sh r02q 2,3,5
