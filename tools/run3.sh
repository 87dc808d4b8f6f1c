set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "512" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_learner.py -x -q -k "cfg3" -s 2>&1 | tail -5
timeout 300 python tools/quick_time.py cfg3 2>&1 | tail -3
for d in 0 1 3; do echo "=== DBG $d"; R2D2_SCAN_DBG=$d timeout 200 python tools/trace_big.py 2>&1 | head -3; done
