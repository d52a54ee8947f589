"""N>1 host logic on CPU: two gloo ranks (127.0.0.1 rendezvous)."""

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["B2F_ROOT"])
from databricks_kubernetes_mlops_poc_b200.sharding import shard_bounds, allgather_merge_moments
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
x = rng.normal(3e4, 9e3, size=(10007, 24))
x[rng.random(x.shape) < 0.01] = np.nan
lo, hi = shard_bounds(len(x), world)[rank]
seg = x[lo:hi]
cnt = (~np.isnan(seg)).sum(0).astype(float)
mean = np.nanmean(seg, 0)
m2 = np.nansum((seg - mean) ** 2, 0)
merged = allgather_merge_moments(np.stack([cnt, mean, m2], axis=1), dist)
ok = (np.allclose(merged[:, 0], (~np.isnan(x)).sum(0)) and np.allclose(merged[:, 1], np.nanmean(x, 0), rtol=1e-12)
      and np.allclose(merged[:, 2] / merged[:, 0], np.nanvar(x, 0), rtol=1e-10))
print(json.dumps({"rank": rank, "world": world, "ok": bool(ok), "rows": hi - lo}))
dist.destroy_process_group()
"""


def _torchrun(args, extra_env=None, timeout=600):
    env = dict(os.environ, B2F_ROOT=ROOT)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29653"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def _json_objects(text: str) -> list:
    """Every JSON object printed on a line of its own -- or glued to another rank's (two processes write to one pipe: a line
    and its newline are not one atomic write)."""
    dec, out = json.JSONDecoder(), []
    for line in text.splitlines():
        i = line.find("{")
        while 0 <= i < len(line):
            try:
                obj, end = dec.raw_decode(line, i)
            except json.JSONDecodeError:
                break
            out.append(obj)
            i = line.find("{", end)
    return out


def test_two_rank_moments_merge(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_objects(r.stdout)
    assert sorted(l["rank"] for l in lines) == [0, 1]
    assert all(l["ok"] and l["world"] == 2 for l in lines)
    assert sum(l["rows"] for l in lines) == 10007


def test_shard_bounds_and_round_robin():
    from databricks_kubernetes_mlops_poc_b200.sharding import round_robin_batches, shard_bounds

    for n in (0, 1, 7, 65536, 10_000_000):
        for k in (1, 2, 3, 8):
            b = shard_bounds(n, k)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    plan = list(round_robin_batches(10_000_000, 65536, 8))
    assert len(plan) == 153 and plan[-1][2] == 10_000_000
    assert [p[0] for p in plan[:10]] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    per_gpu = np.bincount([p[0] for p in plan], minlength=8)
    assert per_gpu.max() - per_gpu.min() <= 1


def test_reference_arm_under_torchrun(tmp_path):
    """`bench.py --impl reference` launched like the driver launches it for N=2: rank 0 prints ONE JSON
    line, rank 1 exits 0 without work."""
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--model", "rf100d6"],
                  extra_env={"B2F_BENCH_CACHE": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_objects(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0
