mkdir -p gpurun_out
python bench.py --gpus 1 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'])
print(json.dumps(d['envelope']['stale_hint'])[:600])
print({k:v['e2e_alignments_per_s'] for k,v in d['configs'].items()})
PY
python -m pytest tests -x -q -m gpu -k "variants or headline or hint or boundary" 2>&1 | tail -3
