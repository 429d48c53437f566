python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo rc=$?
tail -3 gpurun_out/r02_bench_n2.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print(l['e2e']['value']); print(l.get('e2e_inlib')); print(l['per_rank'])
"
python -m pytest tests/test_gpu_parity.py -q -m gpu -k multi_gpu 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
