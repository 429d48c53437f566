#!/bin/bash
# Reproduces the artefacts summarised under profiles/ (run on a B200 box, e.g. `gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh r02'`;
# outputs -> gpurun_out/, then `python tools/ncu_summarize.py gpurun_out/<tag>_step_full.ncu-rep gpurun_out/<tag>_launches_bench.csv <tag>`
# where ncu is installed).  Numbers printed under ncu are never bench values.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
# the two bench lines (own arm, reference arm)
python bench.py --gpus 1 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --gpus 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of a short bench run: the kernels' SHARES of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
# one resident step (pack, bound, wavefront, traceback, CIGAR text, compact), full sections + source view
ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|wfa_bound|wfa_quad|wfa_traceback|cigar_text|cigar_compact" \
    --launch-skip 6 -c 6 -f -o $out/${tag}_step_full python tools/perf_probe.py 8192 10000 0.05 3000 1 1 > $out/${tag}_step_ncu.log 2>&1
# the adaptive band: wavefront kernel + traceback kernel of one chunk
ncu --set full --clock-control none --import-source on -k regex:"wfa_band" --launch-skip 4 -c 2 -f -o $out/${tag}_band_full \
    python tools/band_probe.py 8192 10000 0.05 0.05 2000 25 512 1 > $out/${tag}_band_ncu.log 2>&1
tail -c 400 $out/${tag}_bench.json; echo
