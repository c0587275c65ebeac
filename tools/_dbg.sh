mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -k "parity_case and (mini or far_field or culling or only_2x or plane-)" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cluster_classify" -s 8 -c 1 -o gpurun_out/r02o_cfg5_classify3 -f python tools/run_config_once.py 5 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_cluster_copies" -s 2 -c 1 -o gpurun_out/r02o_cfg3_copies -f python tools/run_config_once.py 3 > /dev/null 2>&1
ls -la gpurun_out/r02o*
timeout 300 bash tools/gpu_cfgs.sh r02o 2,5
