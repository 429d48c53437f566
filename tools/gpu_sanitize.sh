#!/bin/bash
# compute-sanitizer passes over every kernel family (tools/sanitize_target.py); summaries -> gpurun_out/
# usage: tools/gpu_sanitize.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
for tool in memcheck racecheck; do
  for fam in warp cta prebound quad_pairs one_diag score banded banded_one banded_fused large ascii redispatch workers; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $fam > $out/${tag}_san_${tool}_${fam}.log 2>&1
    echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|mismatches' $out/${tag}_san_${tool}_${fam}.log | tr '\n' ' ')"
  done
done | tee $out/${tag}_sanitizer_summary.txt
