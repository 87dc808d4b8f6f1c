set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "512" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_learner.py -x -q -k "cfg3" -s 2>&1 | tail -15
timeout 300 python tools/quick_time.py cfg3 2>&1 | tail -3
R2D2_SCAN_L2XCHG=0 timeout 300 python tools/quick_time.py cfg3 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg3.csv python tools/quick_time.py cfg3 > gpurun_out/ncu_cfg3.log 2>&1; tail -2 gpurun_out/ncu_cfg3.log
