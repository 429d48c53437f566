#!/bin/bash
# bench.py on N GPUs of one box, launched the way the driver does (`gpurun --gpus N -- 'bash tools/gpu_bench_n.sh N'`);
# the JSON line -> gpurun_out/bench_n<N>.json, a short digest on stdout.
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d.get('e2e_inlib'))
for r in d['per_rank']: print(r)
"
grep -v '^\*\|OMP_NUM' gpurun_out/bench_n$N.err | tail -3
