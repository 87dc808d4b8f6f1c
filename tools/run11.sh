mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_step'], d['roofline']['bptt_kernel']['us_per_step'], d['cpu_baseline']['value'], d['other_configs'])"
timeout 600 python bench.py --config replay > gpurun_out/bench_replay.json 2> gpurun_out/bench_replay.err; python -c "
import json; d=json.load(open('gpurun_out/bench_replay.json')); print({k: d[k] for k in ('value','sample_gather_us_per_batch','update_us_per_batch','indices_bit_exact_vs_c_tree','indices_bit_exact_after_update')})"
