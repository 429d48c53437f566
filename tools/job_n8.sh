nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo rc=$?
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02_bench_n8.err | tail -5
python -c "
import json
l=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', l['e2e']['value']); print(l.get('e2e_inlib'))
for r in l['per_rank']: print(r)
"
