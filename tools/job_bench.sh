python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo rc=$?
tail -5 gpurun_out/r02_bench_full.err
python -c "
import json
l=json.load(open('gpurun_out/r02_bench_full.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','pending_after_timed_pass')}); print('e2e', l['e2e']['value'])
print(l['roofline']); print(l['reference_gpu']); print(l.get('e2e_pageable')); print(l['envelope']); print(l.get('cpu_baseline'))
for k,v in l.get('configs',{}).items(): print(k, v)
"
