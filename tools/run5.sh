mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -5 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 python bench.py --config replay > gpurun_out/bench_replay.json 2> gpurun_out/bench_replay.err; tail -3 gpurun_out/bench_replay.err; cat gpurun_out/bench_replay.json
