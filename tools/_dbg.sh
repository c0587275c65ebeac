timeout 200 python -m pytest tests/test_clusters_gpu.py -x -q 2>&1 | tail -6
