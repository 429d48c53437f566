"""N > 1 host logic on the CPU: world_size-2 gloo run of bench.py's sharding / max-over-ranks
reduction, and the library's chunk planning over several GPUs.  No collective exists on the data
path (pairs are independent), so this is all there is to the multi-GPU protocol."""
import ctypes as C
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gloo_world2_sharding_and_max_reduction(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, os.path.join({ROOT!r}, "wfa-gpu_b200", "python"))
        import torch.distributed as dist
        import bench, wfagpu
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        a = wfagpu.Aligner()
        a.add_synthetic(bench.shard_seed(rank), 4, 200, 0.05)
        first = a.pair(0)[0]
        t_max, w_max = bench.reduce_max([1.0 + rank, 10.0 - rank])
        v = bench.aggregate_value(100, world, 3, t_max)
        # one file per rank: two ranks printing to the same pipe can interleave inside a line
        with open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.json"), "w") as f:
            json.dump(dict(rank=rank, world=world, first=first, t_max=t_max, w_max=w_max, v=v), f)
        dist.destroy_process_group()
    """))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), str(script)]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert pr.returncode == 0, pr.stderr[-2000:]
    import json
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in (0, 1)]
    assert sorted(r["rank"] for r in rows) == [0, 1]
    assert rows[0]["first"] != rows[1]["first"]                  # disjoint shards
    for r in rows:
        assert r["world"] == 2 and r["t_max"] == 2.0 and r["w_max"] == 10.0
        assert r["v"] == 100 * 2 * 3 / 2.0                       # whole-job pairs / slowest rank


def test_reference_arm_runs_on_rank0_only(tmp_path):
    env = dict(os.environ, WFAGPU_BENCH_REF_PAIRS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference",
           "--gpus", "2", "--steps", "1", "--warmup", "0"]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert pr.returncode == 0, pr.stderr[-2000:]
    lines = [l for l in pr.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_chunk_planning(lib):
    def plan(n, batch, ndev, span):
        c, k = C.c_size_t(), C.c_size_t()
        lib.wfagpu_plan_chunks(n, batch, ndev, span, C.byref(c), C.byref(k))
        return c.value, k.value
    assert plan(1000, 100, 1, 1000 * 300) == (100, 10)
    assert plan(1000, 0, 1, 1000 * 300) == (1000, 1)            # batch 0 = one batch
    assert plan(1000, 5000, 1, 1000 * 300) == (1000, 1)
    c, k = plan(100000, 100000, 8, 100000 * 20000)
    assert k >= 16 and c * k >= 100000                           # >= 2 chunks per GPU
    c, k = plan(1000000, 1000000, 1, 1000000 * 20000)            # 20 GB of ASCII: stays below 32-bit offsets
    assert c * 20000 < (1 << 32)
    assert plan(0, 10, 4, 0) == (0, 0)
    # every pair is covered exactly once
    for n, b, d in ((7, 3, 2), (100, 7, 8), (33, 33, 4)):
        c, k = plan(n, b, d, n * 400)
        assert (k - 1) * c < n <= k * c
