for i in 1 2; do timeout 300 bash tools/gpu_cfgs.sh r02w 3,5; done
