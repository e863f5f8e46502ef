"""N > 1 host logic on CPU (world size 2, gloo): the window -> rank assignment of bench.py's strong-scaling mode covers every
window exactly once, and the max-over-ranks reduction used for the reported time works without a GPU."""
import os
import socket
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent('''
    import os, sys, json
    import torch, torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    total = int(sys.argv[1])
    mine = list(range(rank, total, world))            # bench.py: window w -> rank w mod world (SURVEY 8 e)
    seeds = [rank + w * world for w in range(len(mine))]
    assert seeds == mine
    t = torch.zeros(total, dtype=torch.int64)
    t[mine] = 1
    dist.all_reduce(t)                                 # every window owned exactly once
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # time = max over ranks
    if rank == 0:
        print(json.dumps({"covered": bool((t == 1).all()), "max_ms": float(ms), "n0": len(mine)}))
    dist.destroy_process_group()
''')


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_round_robin_sharding_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script), "121"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    import json

    j = json.loads(line)
    assert j == {"covered": True, "max_ms": 11.0, "n0": 61}
