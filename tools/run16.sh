timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -2
R2D2_OVERLAP_INPUTS=0 timeout 300 python tools/quick_time.py cfg3 cfg2 | tail -2
R2D2_OVERLAP_INPUTS=1 timeout 300 python tools/quick_time.py cfg3 cfg2 | tail -2
R2D2_OVERLAP_INPUTS=0 timeout 300 python tools/quick_time.py cfg3 | tail -1
R2D2_OVERLAP_INPUTS=1 timeout 300 python tools/quick_time.py cfg3 | tail -1
