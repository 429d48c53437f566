set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
tools/gpu_sanitize.sh r02a
python tools/ref_gpu_probe.py 2048 10000 0.05 3000 1 > gpurun_out/r02a_refgpu_10k.json 2> gpurun_out/r02a_refgpu_10k.err
cat gpurun_out/r02a_refgpu_10k.json
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cat gpurun_out/r02a_bench.json | cut -c1-600
