timeout 400 python -m pytest tests -m gpu -x -q -k "parity_case or properties or sharded or culling or headline or config3 or mailbox" 2>&1 | tail -3
timeout 300 bash tools/gpu_cfgs.sh r03b 1,2,3
