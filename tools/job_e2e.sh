for fc in 1 64 256 512; do WFAGPU_FIRST_CHUNK=$fc python tools/e2e_probe.py 8192 10000 0.05 3000 1 4096 5 | python -c "
import sys,json; l=json.loads(sys.stdin.read()); print('first', l['env'].get('WFAGPU_FIRST_CHUNK'), 'batch', l['batch'], 'ms', l['wall_ms'], 'aln/s', l['pairs_per_s'])"; done
WFAGPU_RAMP=0 python tools/e2e_probe.py 8192 10000 0.05 3000 1 4096 5 | cut -c1-200
echo ---- first=1
WFAGPU_FIRST_CHUNK=1 WFAGPU_VERBOSE=1 WFAGPU_TRACE=1 python tools/e2e_probe.py 8192 10000 0.05 3000 1 4096 1 2>&1 | grep -v "pass:" | tail -12 | cut -c1-250
echo ---- ramp=0
WFAGPU_RAMP=0 WFAGPU_VERBOSE=1 WFAGPU_TRACE=1 python tools/e2e_probe.py 8192 10000 0.05 3000 1 4096 1 2>&1 | grep -v "pass:" | tail -8 | cut -c1-250
