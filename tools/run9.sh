timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "512" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_learner.py -x -q -k "cfg3 or resume" -s 2>&1 | tail -4
timeout 300 python tools/time_scan.py
REP=2 S=160 timeout 300 python tools/time_scan.py | grep bwd
R2D2_SCAN_BWD_PP=0 timeout 300 python tools/time_scan.py | grep bwd
timeout 300 python tools/quick_time.py cfg3 | tail -1
