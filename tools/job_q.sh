tools/gpu_sanitize.sh r02b
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|wfa_bound|wfa_quad|wfa_traceback|cigar_text|cigar_compact" --launch-skip 6 -c 6 -f -o gpurun_out/r02_step_full python tools/perf_probe.py 8192 10000 0.05 3000 1 1 > gpurun_out/r02_step_ncu.log 2>&1; tail -2 gpurun_out/r02_step_ncu.log
