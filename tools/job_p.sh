python tools/e2e_probe.py 100000 10000 0.03 3000 1 100000 2 2>&1 | tail -1 | cut -c1-300
python - <<'PY'
import sys, time, json
sys.path.insert(0, "wfa-gpu_b200/python")
import wfagpu
for env in ("default",):
    a = wfagpu.Aligner(); a.add_synthetic(0xB2000044, 100000, 10000, 0.01, 0.05); a.initialize_parameters(2,3,1)
    a.options.max_error = 3000; a.options.compute_cigar = True
    a.align(); ts=[]
    for _ in range(2):
        a.reset_results(); t0=time.perf_counter(); a.align(); ts.append(time.perf_counter()-t0)
    print("mix 1-5%", [round(t*1e3,1) for t in ts], round(100000/min(ts)))
PY
WFAGPU_NO_BOUND_ORDER=1 python - <<'PY'
import sys, time, json
sys.path.insert(0, "wfa-gpu_b200/python")
import wfagpu
a = wfagpu.Aligner(); a.add_synthetic(0xB2000044, 100000, 10000, 0.01, 0.05); a.initialize_parameters(2,3,1)
a.options.max_error = 3000; a.options.compute_cigar = True
a.align(); ts=[]
for _ in range(2):
    a.reset_results(); t0=time.perf_counter(); a.align(); ts.append(time.perf_counter()-t0)
print("mix 1-5% no order", [round(t*1e3,1) for t in ts], round(100000/min(ts)))
PY
python tools/perf_probe.py 8192 10000 0.05 3000 1 4 2>&1 | tail -1 | cut -c1-330
python tools/e2e_probe.py 8192 10000 0.05 3000 1 8192 5 2>&1 | tail -1 | cut -c1-300
python -m pytest tests -x -q -m gpu -k "variants or headline or config" 2>&1 | tail -3
