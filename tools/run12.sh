mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_learner.py -x -q 2>&1 | tail -3
timeout 300 python tools/time_scan.py
timeout 300 python tools/quick_time.py cfg3 cfg2 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --episodes 48 > gpurun_out/ncu_ll.log 2>&1; tail -1 gpurun_out/ncu_ll.log
