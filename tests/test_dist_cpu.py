"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard arithmetic and the replay-sample gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from onitama_alphazero_b200.sharding import gather_replay, shard_range
    first, count = shard_range(1001, rank, world)
    m = 3 + 4 * rank  # ragged sample counts
    g = torch.Generator().manual_seed(100 + rank)
    planes = (torch.rand((m, 21, 5, 5), generator=g) > 0.5).float()
    pi = torch.rand((m, 2, 25), generator=g)
    z = torch.full((m,), float(rank) - 0.5)
    out = gather_replay(planes, pi, z, dst=0)
    if rank == 0:
        q.put((first, count, out[0].numpy(), out[1].numpy(), out[2].numpy()))
    else:
        assert out is None
        q.put((first, count, None, None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers():
    from onitama_alphazero_b200.sharding import shard_range
    for n in (1, 7, 1000, 1 << 20):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_gather_replay_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    assert res[0][:2] == (0, 500) and res[1][:2] == (500, 501)
    planes, pi, z = res[0][2], res[0][3], res[0][4]
    assert planes.shape == (3 + 7, 21, 5, 5) and pi.shape == (10, 2, 25)
    assert z.tolist() == [-0.5] * 3 + [0.5] * 7
    for rank, sl in ((0, slice(0, 3)), (1, slice(3, 10))):
        g = torch.Generator().manual_seed(100 + rank)
        m = 3 + 4 * rank
        want_planes = (torch.rand((m, 21, 5, 5), generator=g) > 0.5).float().numpy()
        want_pi = torch.rand((m, 2, 25), generator=g).numpy()
        assert np.array_equal(planes[sl], want_planes) and np.array_equal(pi[sl], want_pi)


def test_reference_arm_under_torchrun_prints_one_json_line():
    """bench.py --impl reference launched the way the driver launches N > 1: rank 0 alone measures and prints, the other
    rank exits 0 without work, stdout carries exactly one JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29593", os.path.join(root, "bench.py"), "--gpus", "2", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"] == "env_steps_per_sec"
    # the arm runs the config it prints: all 1 048 576 games, one lockstep step of them per timed step
    assert d["config"]["games_per_gpu"] == 1 << 20 and d["cpu_baseline"]["sample"].startswith("1048576 games x 1 lockstep step")
