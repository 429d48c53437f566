WFAGPU_VERBOSE=1 python bench.py --quick --steps 3 --warmup 2 --no-cpu-baseline 2> gpurun_out/o.err | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print(l['value'], l['e2e']['value'], l['per_rank'])"
grep "chunk from" gpurun_out/o.err | tail -6 | cut -c1-200
OMP_NUM_THREADS=1 python tools/e2e_probe.py 8192 10000 0.05 3000 1 8192 5 | cut -c1-200
