python tools/cigar_validity_probe.py 4096 10000 0.05 3000 2048
python tools/cigar_validity_probe.py 20000 1000 0.10 400 10000
python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_vs_reference_gpu.py -x -q -m gpu 2>&1 | tail -8
