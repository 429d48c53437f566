python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo rc=$?
tail -5 gpurun_out/r02_bench_full.err
python -c "
import json
l=json.load(open('gpurun_out/r02_bench_full.json'))
print({k:l[k] for k in ('value','ms_per_step','e2e','gpu_launches','pending_after_timed_pass')})
print(l['roofline']); print(l['reference_gpu']); print(l.get('e2e_pageable')); print(l['envelope']); print(l.get('cpu_baseline'))
for k,v in l.get('configs',{}).items(): print(k, v)
print(l['per_rank'])
"
