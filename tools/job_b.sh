set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_vs_reference_gpu.py -x -q -m gpu 2>&1 | tail -15
for t in 0 64 96 128 160 192 256; do
  WFAGPU_THREADS=$t python tools/perf_probe.py 8192 10000 0.05 3000 1 3
done 2>&1 | tee gpurun_out/r02b_quad_threads.jsonl
WFAGPU_NO_QUAD=1 python tools/perf_probe.py 8192 10000 0.05 3000 1 3 | tee -a gpurun_out/r02b_quad_threads.jsonl
python tools/perf_probe.py 8192 10000 0.05 3000 0 3 | tee -a gpurun_out/r02b_quad_threads.jsonl
WFAGPU_NO_QUAD=1 python tools/perf_probe.py 8192 10000 0.05 3000 0 3 | tee -a gpurun_out/r02b_quad_threads.jsonl
python tools/perf_probe.py 50000 1000 0.10 400 1 3 | tee -a gpurun_out/r02b_quad_threads.jsonl
WFAGPU_NO_QUAD=1 python tools/perf_probe.py 50000 1000 0.10 400 1 3 | tee -a gpurun_out/r02b_quad_threads.jsonl
