#!/usr/bin/env python3
"""bench.py -- headline benchmark of the gap-affine WFA hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our CUDA path)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU WFA, host cores)

Workload (BASELINE.json metric, SURVEY.md 8(d) cfg 4 headline sub-run): synthetic
10 kbp pairs, 5 % error, penalties x=2,o=3,e=1, `-e 3000`, exact, with CIGAR.
A step = one pass of the hot path (pack, score-bound, wavefront, traceback and CIGAR-text
kernels) over one batch of PAIRS_PER_GPU pairs per GPU.  Pairs shard across GPUs with no collective
(weak scaling: per-GPU batch fixed).

  value : alignments/s over all GPUs, batch resident in HBM, device time from CUDA
          events recorded on the launching stream, max over ranks.
  e2e   : same metric through the public C API (wfagpu_align) from page-locked host
          buffers: H2D of the ASCII, kernels, D2H of results + op streams and CIGAR
          text generation are all inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "wfa-gpu_b200", "python"))

METRIC = "alignments_per_s_10kbp_5pct_cigar"
UNIT = "alignments/s"
LENGTH, ERR, PEN, MAX_ERROR = 10000, 0.05, (2, 3, 1), 3000
PAIRS_PER_GPU = int(os.environ.get("WFAGPU_BENCH_PAIRS", 8192))
CPU_SAMPLE = int(os.environ.get("WFAGPU_BENCH_CPU_SAMPLE", 1536))
REF_STEP_PAIRS = int(os.environ.get("WFAGPU_BENCH_REF_PAIRS", 768))
SEED = 0xB2000004


def config(extra=None):
    c = {"workload": "cfg4-headline: 10 kbp pairs, 5% error, x=2,o=3,e=1, -e 3000, exact, CIGAR",
         "pairs_per_gpu_per_step": PAIRS_PER_GPU, "length": LENGTH, "error_rate": ERR,
         "penalties": list(PEN), "max_error": MAX_ERROR,
         "l2_policy": "inputs+arenas larger than L2 (164 MB ASCII + >10 GB ring-snapshot arena per step)"}
    if extra:
        c.update(extra)
    return c


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line).

    Sampled through NVML inside this process (nvidia_ml_py): starting an nvidia-smi process every
    200 ms takes the driver lock for tens of milliseconds and showed up as idle gaps between the
    kernels of a step.  Falls back to nvidia-smi when NVML cannot be loaded."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []          # (sm_mhz, sm_max_mhz, [reason names])
        self.stop_flag = threading.Event()
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu])
                except Exception:
                    idx = gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        masks = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        self.rows.append((int(sm), int(mx), [self.NAMES[i] for i in range(4) if bits & masks[i]]))

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [x.strip() for x in out.split(",")]
        if len(r) >= 6 and r[0].isdigit():
            self.rows.append((int(r[0]), int(r[1]) if r[1].isdigit() else 0,
                              [self.NAMES[i] for i in range(4) if r[2 + i].lower().startswith("active")]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.1 if self.nvml else 0.5)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows if r[1]]
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def shard_seed(rank):
    """Every rank aligns its own deterministic slice of the synthetic workload."""
    return SEED + 7919 * rank


def reduce_max(values, device=None):
    """Max over ranks of a list of floats (device time, wall time): the slowest rank sets the pace."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_value(pairs_per_gpu, world, steps, t_max):
    """Whole-job throughput: all pairs of all ranks over the slowest rank's time."""
    return pairs_per_gpu * world * steps / t_max


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def algorithmic_bytes(a, n_ops_total):
    """ASCII read by pack + packed write + one packed read + metadata + result + bt chain."""
    total = 0
    for i in range(a.num_pairs):
        m = a.s.sequences_metadata[i]
        pl, tl = m.pattern_len, m.text_len
        total += (pl + tl) + 2 * ((pl + 3) // 4 + (tl + 3) // 4) + 48 + 20
    return total + 8 * ((n_ops_total + 15) // 16)


def cells_of_scores(lib, wfagpu, scores):
    import ctypes as C
    x, o, e = PEN
    md = MAX_ERROR * (max(x, o + e) + 1) + 16
    tab = (wfagpu.Step * (md + 1))()
    units = C.c_uint64()
    d_end = lib.wfagpu_build_step_table(x, o, e, MAX_ERROR, md, 0, tab, C.byref(units))
    cum = [0] * (d_end + 1)
    run = 0
    for d in range(d_end):
        if tab[d].kind == 2:
            run += 2 * tab[d].n + 1
        elif tab[d].kind == 1:
            run += 2 * tab[d].n + 1
        cum[d] = run
    return sum(cum[min(s, d_end - 1)] for s in scores)


def measured_cells(gpu):
    """Cells the wavefront kernel really computes for one step of this workload: the same batch run
    once more in a child process with the kernel's per-pair cell counter switched on (untimed)."""
    try:
        env = dict(os.environ, WFAGPU_COUNT_CELLS="1", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(gpu)))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "perf_probe.py"), str(PAIRS_PER_GPU), str(LENGTH),
                              str(ERR), str(MAX_ERROR), "1", "1"], env=env, capture_output=True, text=True, timeout=300).stdout
        return int(json.loads(out.strip().splitlines()[-1])["cells"])
    except Exception:
        return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    import wfagpu

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = wfagpu.load()
    wfagpu.set_devices(str(local))

    a = wfagpu.Aligner()
    a.add_synthetic(shard_seed(rank), PAIRS_PER_GPU, LENGTH, ERR, ERR)
    assert a.initialize_parameters(*PEN)
    a.options.max_error = MAX_ERROR
    a.options.compute_cigar = True
    a.set_batch_size(max(1, PAIRS_PER_GPU // 2))          # two chunks: the second one's copies overlap the first one's kernels
    gcells_total = sum(a.s.sequences_metadata[i].pattern_len * a.s.sequences_metadata[i].text_len
                       for i in range(a.num_pairs))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident hot path: `value` ----------------------
    rb = wfagpu.ResidentBatch(a, device=local, slot=0)
    rb.upload()
    plan = rb.plan(cigar=True)
    for _ in range(args.warmup):
        rb.align(plan)
        rb.wait()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    align_ms = 0.0
    wf_ms = 0.0
    for _ in range(args.steps):
        rb.align(plan)
        mp, ma = rb.wait()
        dev_ms += mp + ma
        align_ms += ma
        wf_ms += rb.stats()["ms_wavefront"]        # CUDA events around the wavefront kernel on its stream
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.summary()
    st = rb.stats()
    launches_per_step = st["launches"]
    out, ops, used = rb.download()
    scores = [out[i].distance for i in range(rb.n)]
    n_ops_total = sum(out[i].n_ops for i in range(rb.n))
    assert all(out[i].status & 1 for i in range(rb.n)), "unfinished pairs in the benchmark batch"

    t_max, wall_max = reduce_max([dev_ms / 1e3, wall], device="cuda")
    total_pairs = PAIRS_PER_GPU * world * args.steps
    value = aggregate_value(PAIRS_PER_GPU, world, args.steps, t_max)

    # ---------------- end to end through the public C API: `e2e` -------------
    a.pin_host_buffers()
    for _ in range(max(1, min(args.warmup, 2))):
        a.reset_results()
        a.align()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a.reset_results()
        a.align()
    barrier()
    e2e_s = time.perf_counter() - t0
    rs = a.run_stats()
    (t_e2e_max,) = reduce_max([e2e_s], device="cuda")
    e2e_value = total_pairs / t_e2e_max
    # the e2e answer must be the resident answer
    assert [a.error(i) for i in range(0, a.num_pairs, 97)] == scores[::97]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = hbm_peak()
    alg_bytes = algorithmic_bytes(a, n_ops_total)
    k_ms = (wf_ms or align_ms) / args.steps      # dominant kernel: wfa_exact_kernel, its own launch duration
    achieved = alg_bytes / (k_ms / 1e3) / 1e9
    cells_unpruned = cells_of_scores(lib, wfagpu, scores)           # what the reference's kernels compute for these scores
    cells = measured_cells(local) or cells_unpruned                 # cells the wavefront kernel computed in one step
    sm = rb.sm_count()
    clk = (clocks["sm_mhz"] or 1965) * 1e6
    int_peak = sm * 128 * clk
    roofline = {"bound": "hbm", "kernel": "wfa_exact_kernel<cta>", "achieved": round(achieved, 2), "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 6),
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(k_ms, 3),
                "step_kernels_ms": round(align_ms / args.steps, 3), "traffic": None,
                "note": "compulsory HBM traffic is ~31 KB/pair: the path is issue-bound, see roofline_int"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = int(json.load(open(traffic_file))["dram_bytes_per_pair"] * PAIRS_PER_GPU)
            roofline["traffic_note"] = "ncu dram bytes per pair (profiles/traffic.json) x pairs per launch"
        except Exception:
            pass
    roofline_int = {"bound": "issue", "cells_per_launch": cells, "cells_without_pruning": cells_unpruned,
                    "nominal_instr_per_cell": 64,
                    "achieved": round(cells * 64 / (k_ms / 1e3) / 1e12, 3), "peak": round(int_peak / 1e12, 3),
                    "unit": "T thread-instr/s", "frac": round(cells * 64 / (k_ms / 1e3) / int_peak, 4),
                    "gcells_per_s": round(cells / (k_ms / 1e3) / 1e9, 3),
                    # SURVEY 8(d) defines the work unit with the reference's growth model (cells_without_pruning):
                    # the same time against that work -- above 1 means fewer cells than the reference computes
                    "frac_of_reference_work": round(cells_unpruned * 64 / (k_ms / 1e3) / int_peak, 4)}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(a, CPU_SAMPLE)

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t_max * 1e3 / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": config({"parallelism": f"pairs sharded over {world} GPU(s), no collective"}),
        "gcups": round(gcells_total * world * args.steps / t_max / 1e9, 1),
        "wall_ms_per_step": round(wall_max * 1e3 / args.steps, 3),
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT,
                "h2d_bytes_per_step": int(rs["h2d_bytes"]), "d2h_bytes_per_step": int(rs["d2h_bytes"]),
                "gcups": round(gcells_total * world * args.steps / t_e2e_max / 1e9, 1),
                "api": "wfagpu_align (page-locked host buffers; H2D, kernels, CIGAR text printed on the GPU, D2H, copy into results[i])"},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks, "roofline": roofline, "roofline_int": roofline_int,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a, sample):
    """The reference's CPU path (WFA2-lib v2.3 through utils/wfa_cpu.c, compiled into oracle/_ref)
    on the first `sample` pairs of the workload, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import RefCPU, Oracle
    n = min(sample, a.num_pairs)
    pairs = [a.pair(i) for i in range(n)]
    if RefCPU.available():
        r = RefCPU()
        threads = len(os.sched_getaffinity(0))          # all host cores (torchrun pins OMP_NUM_THREADS=1)
        t0 = time.perf_counter()
        errs, _ = r.align_batch([p for p, _ in pairs], [t for _, t in pairs], *PEN, cigar=True, threads=threads)
        dt = time.perf_counter() - t0
        for i in range(0, n, 37):
            assert errs[i] == a.error(i), "CPU reference and GPU scores differ"
        return {"value": round(n / dt, 1), "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"first {n} pairs of the workload, CIGAR, wavefront_memory_low, OpenMP static",
                "seconds": round(dt, 2)}
    o = Oracle()
    n = min(n, 24)
    t0 = time.perf_counter()
    for p, t in pairs[:n]:
        o.align(p, t, *PEN, MAX_ERROR)
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 2), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n} pairs, scalar restatement", "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    import wfagpu
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import RefCPU, Oracle
    a = wfagpu.Aligner()
    a.add_synthetic(SEED, REF_STEP_PAIRS, LENGTH, ERR, ERR)
    pairs = [a.pair(i) for i in range(a.num_pairs)]
    P, T = [p for p, _ in pairs], [t for _, t in pairs]
    if RefCPU.available():
        r = RefCPU()
        threads, kind = len(os.sched_getaffinity(0)), "reference"
        step = lambda: r.align_batch(P, T, *PEN, cigar=True, threads=threads)
        n = len(P)
    else:
        o = Oracle()
        threads, kind, n = 1, "port", 8
        step = lambda: [o.align(p, t, *PEN, MAX_ERROR) for p, t in pairs[:n]]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    gc = sum(len(p) * len(t) for p, t in pairs[:n]) * args.steps / dt / 1e9
    sample = f"{n} pairs per step (bounded sample of the workload), CIGAR, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 1), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3 / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config({"pairs_per_step": n}), "gcups": round(gc, 2),
        "cpu_baseline": {"value": round(v, 1), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
