timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k 512 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_learner.py -x -q -k "cfg3" -s 2>&1 | tail -3
timeout 300 python tools/time_scan.py
timeout 300 python tools/quick_time.py cfg3 | tail -1
