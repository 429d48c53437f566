for cfg in "50000 1000 0.10 400" "30000 2000 0.08 600" "20000 3000 0.05 600" "12000 5000 0.05 1000"; do
  set -- $cfg
  python tools/perf_probe.py $1 $2 $3 $4 1 3 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('quad  ', l['pairs'], l['len'], l['err'], 'align', min(l['align_ms']), 'wf', l['wavefront_ms'], 'n_cap', l['n_cap'], 'thr', l['cta_threads'], 'ctas', l['ctas'])"
  WFAGPU_NO_QUAD=1 python tools/perf_probe.py $1 $2 $3 $4 1 3 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('noquad', l['pairs'], l['len'], l['err'], 'align', min(l['align_ms']), 'wf', l['wavefront_ms'], 'n_cap', l['n_cap'], 'thr', l['cta_threads'], 'ctas', l['ctas'])"
done
