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
    sys.path.insert(0, os.getcwd())
    from norma_b200 import workload                   # the SAME functions bench.py shards with (SURVEY 8 e: window w -> rank w mod world)
    mine = workload.windows_of_rank(rank, world, total)
    assert mine == workload.window_ids(rank, world, 25, total)
    assert workload.window_ids(rank, world, 25) == list(range(rank * 25, rank * 25 + 25))
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


def test_sharding_rule_edge_cases():
    from norma_b200 import workload as wl

    for world in (1, 2, 3, 8):
        for total in (0, 1, 5, 8, 120, 121):
            owned = [w for r in range(world) for w in wl.windows_of_rank(r, world, total)]
            assert sorted(owned) == list(range(total))                       # every window exactly once
            sizes = [len(wl.windows_of_rank(r, world, total)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1                               # balanced
    assert wl.windows_of_rank(7, 8, 5) == []                                  # fewer windows than ranks: that rank has nothing (bench.py refuses it)
    assert wl.windows_per_step(8, 25) == 200 and wl.windows_per_step(8, 25, 120) == 120
    import pytest

    with pytest.raises(ValueError):
        wl.windows_of_rank(2, 2, 10)


def test_algorithmic_work_figures_match_survey():
    """36.94 / 87.37 / 2273.77 GF per window; mel 2.880 / 3.456 MB; decode 56 / 97 / 225 / 1601 MB per token."""
    from norma_b200 import synth, workload as wl

    gf = lambda n: wl.encoder_flops(synth.model_config(n)) / 1e9
    assert abs(gf("tiny.en") - 36.94) < 0.01 and abs(gf("base.en") - 87.37) < 0.01 and abs(gf("distil-large-v3") - 2273.77) < 0.01
    assert wl.mel_bytes(synth.model_config("tiny.en")) == 2.88e6 and wl.mel_bytes(synth.model_config("large-v3")) == 3.456e6
    c = synth.model_config("distil-large-v3")
    assert abs((wl.decode_bytes_per_step(c, 0)) / 1e6 - 225) < 1
    assert abs(wl.decode_bytes_per_step(synth.model_config("large-v3"), 0) / 1e6 - 1601) < 1
    assert abs(wl.gemm_flops(c) + wl.attention_flops(c) - wl.encoder_flops(c)) < 1
