mkdir -p gpurun_out
python bench.py --gpus 1 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -c 6000 gpurun_out/r02e_bench.json
tail -5 gpurun_out/r02e_bench.err
