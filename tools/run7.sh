mkdir -p gpurun_out
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --episodes 48 > gpurun_out/ncu_ll.log 2>&1; tail -2 gpurun_out/ncu_ll.log
# full captures of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_scan_fwd_big -s 2 -c 1 -o gpurun_out/r02_scan_fwd_big python tools/time_scan.py > gpurun_out/ncu_fwd.log 2>&1; tail -2 gpurun_out/ncu_fwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_scan_bwd_tc -s 1 -c 1 -o gpurun_out/r02_scan_bwd python tools/time_scan.py > gpurun_out/ncu_bwd.log 2>&1; tail -2 gpurun_out/ncu_bwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:td_elem|td_reduce|gather_batch|tree_sample|tree_update" -s 10 -c 5 -o gpurun_out/r02_td_gather python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --episodes 48 > gpurun_out/ncu_td.log 2>&1; tail -2 gpurun_out/ncu_td.log
ls -la gpurun_out/*.ncu-rep
