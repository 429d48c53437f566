WFAGPU_VERBOSE=1 python tools/cfg5_probe.py 256 2>&1 | grep -v "chunk from" | cut -c1-300 | tail -12
ncu --set full --clock-control none --import-source on -k regex:wfa_exact --launch-skip 1 -c 1 -f -o gpurun_out/r02_large_full python tools/cfg5_probe.py 64 > gpurun_out/r02_large_ncu.log 2>&1; tail -2 gpurun_out/r02_large_ncu.log
