mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/r02_final_bench_n$N.json 2> gpurun_out/r02_final_bench_n$N.err
tail -1 gpurun_out/r02_final_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d.get('e2e_inlib'))
for r in d['per_rank']: print(r)
print({k:v.get('e2e_alignments_per_s') for k,v in d.get('configs',{}).items()})
"
tail -3 gpurun_out/r02_final_bench_n$N.err
