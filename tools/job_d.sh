set -x
python tests/variant_check.py > gpurun_out/r02d_variant_default.log 2>&1; echo rc=$?; tail -4 gpurun_out/r02d_variant_default.log
WFAGPU_FORCE_BOUND=1 python tests/variant_check.py > gpurun_out/r02d_variant0.log 2>&1; echo rc=$?; tail -4 gpurun_out/r02d_variant0.log
for t in 0 128 160 192 224; do
  WFAGPU_THREADS=$t python tools/perf_probe.py 8192 10000 0.05 3000 1 3
done 2>&1 | tee gpurun_out/r02d_quad.jsonl
WFAGPU_NO_QUAD_PAIRS=1 python tools/perf_probe.py 8192 10000 0.05 3000 1 3 | tee -a gpurun_out/r02d_quad.jsonl
python tools/perf_probe.py 8192 10000 0.05 3000 0 3 | tee -a gpurun_out/r02d_quad.jsonl
python tools/perf_probe.py 50000 1000 0.10 400 1 3 | tee -a gpurun_out/r02d_quad.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_vs_reference_gpu.py -x -q -m gpu 2>&1 | tail -8
