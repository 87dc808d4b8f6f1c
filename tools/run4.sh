timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "512" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_learner.py -x -q -k "cfg3" -s 2>&1 | tail -4
for d in 0 1 2 3; do echo "=== DBG $d"; R2D2_SCAN_DBG=$d timeout 200 python tools/time_scan.py 2>&1 | grep bwd; done
REP=2 S=160 timeout 200 python tools/time_scan.py
timeout 300 python tools/quick_time.py cfg3 | tail -1
