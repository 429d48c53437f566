set -x
WFAGPU_FORCE_BOUND=1 python tests/variant_check.py > gpurun_out/r02c_variant0.log 2>&1; echo rc=$?; tail -30 gpurun_out/r02c_variant0.log
ncu --set full --clock-control none --import-source on -k regex:wfa_quad --launch-skip 1 -c 1 -f -o gpurun_out/r02c_quad_full python tools/perf_probe.py 2960 10000 0.05 3000 1 1 > gpurun_out/r02c_ncu.log 2>&1; tail -3 gpurun_out/r02c_ncu.log
