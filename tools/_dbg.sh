timeout 400 python -m pytest tests -m gpu -x -q -k "parity_case or config5 or config3 or culling or properties" 2>&1 | tail -3
timeout 300 bash tools/gpu_cfgs.sh r02z 2,3,5
