"""N > 1 host-side logic on CPU (torch.distributed, gloo, world_size 2; no GPU):

* the shared stream is indexed by the GLOBAL particle id, so the draws each rank generates for its
  own shard, all-gathered, are exactly the draws a single process generates for the whole
  population -- the property the sharded filter's bit-equality with the single-GPU run rests on;
* the rank-reduction helpers of bench.py (max over ranks for times, sum for counts);
* `bench.py --impl reference` under torchrun: rank 0 alone runs and prints the JSON line, the other
  rank exits 0 without work.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SLAMRS_ROOT"])
from oracle import oracle as O

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
SEED, N = 0x5EED5A11, 64
S = N // world
ok = True
for step in (0, 3, 2**33 + 1):
    mine = torch.from_numpy(O.motion_normals(SEED, step, rank * S, S).copy())        # this rank's shard
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    whole = O.motion_normals(SEED, step, 0, N)                                          # single-process draw
    ok &= bool(np.array_equal(torch.cat(parts).numpy().view(np.uint64), whole.view(np.uint64)))
    u = torch.tensor([O.resample_uniform(SEED, step)], dtype=torch.float64)
    us = [torch.empty_like(u) for _ in range(world)]
    dist.all_gather(us, u)
    ok &= all(float(x) == float(u) for x in us)                                         # replicated, identical
# bench.py's reductions: max over ranks for times, sum over ranks for counts
t = torch.tensor([1.0 + rank, 5.0 - rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok &= t.tolist() == [float(world), 5.0]
c = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
dist.all_reduce(c, op=dist.ReduceOp.SUM)
ok &= c.item() == 10.0 * world * (world + 1) / 2
# contiguous shards: rank g owns [g*S, (g+1)*S) before and after resampling
owners = np.arange(N) // S
ok &= bool(np.all(owners[rank * S:(rank + 1) * S] == rank))
print(json.dumps({"rank": rank, "ok": bool(ok)}), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _torchrun(args, env_extra=None, timeout=240):
    # build both libraries here, once, so that the two ranks never race to (re)build them
    from oracle import oracle as O
    from slamrs_b200 import _lib
    O.lib(); _lib.load()
    env = dict(os.environ)
    env.update({"SLAMRS_ROOT": ROOT, "OMP_NUM_THREADS": "1", "CUDA_VISIBLE_DEVICES": ""})
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_sharded_stream_and_reductions_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = _torchrun([str(script)])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    import re
    # the two ranks share stdout and their lines may interleave: pick the records out individually
    lines = [json.loads(m) for m in re.findall(r'\{"rank": \d+, "ok": (?:true|false)\}', out.stdout)]
    assert sorted(l["rank"] for l in lines) == [0, 1] and all(l["ok"] for l in lines)


def test_reference_arm_under_torchrun_world2():
    out = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                     "--workload", "c1", "--cpu-particles", "4"])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly rank 0 prints"                      # the other rank exits 0 without work
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
