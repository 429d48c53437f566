python tools/cigar_validity_probe.py 4096 10000 0.05 3000 2048
WFAGPU_NO_QUAD_PAIRS=1 python tools/cigar_validity_probe.py 20000 1000 0.10 400 10000
for t in 96 128 160; do
  WFAGPU_THREADS=$t python tools/perf_probe.py 8192 10000 0.05 3000 1 3
  WFAGPU_NO_QUAD_PAIRS=1 WFAGPU_THREADS=$t python tools/perf_probe.py 8192 10000 0.05 3000 1 3
done 2>&1 | tee gpurun_out/r02g_quad.jsonl
python tools/perf_probe.py 8192 10000 0.05 3000 0 3 | tee -a gpurun_out/r02g_quad.jsonl
python tools/perf_probe.py 50000 1000 0.10 400 1 3 | tee -a gpurun_out/r02g_quad.jsonl
WFAGPU_NO_QUAD_PAIRS=1 python tools/perf_probe.py 50000 1000 0.10 400 1 3 | tee -a gpurun_out/r02g_quad.jsonl
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_vs_reference_gpu.py -x -q -m gpu 2>&1 | tail -4
