mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_step'], d['roofline']['bptt_kernel']['us_per_step'], d['cpu_baseline']['value'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('reference arm', d['value'], d['steps_timed'], d['cpu_baseline']['cores'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --episodes 48 > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_scan_bwd_tc -s 1 -c 2 -o gpurun_out/r02_scan_bwd python tools/time_scan.py > gpurun_out/ncu_bwd.log 2>&1; tail -1 gpurun_out/ncu_bwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_scan_fwd_big -s 2 -c 1 -o gpurun_out/r02_scan_fwd_big python tools/time_scan.py > gpurun_out/ncu_fwd.log 2>&1; tail -1 gpurun_out/ncu_fwd.log
