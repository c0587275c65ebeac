timeout 300 python -m pytest tests/test_properties_gpu.py -x -q -k "mailbox or sharded" 2>&1 | tail -4
