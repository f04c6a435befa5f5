"""One rank of the 2-GPU test of the replay gather behind the C ABI (tests/test_gpu_parity.py::test_gather_samples_two_gpus):
onb_comm_* / onb_gather_samples against sharding.gather_replay (torch.distributed, NCCL) on the same samples.
    python tests/dist_gather_worker.py RANK WORLD PORT IDFILE"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank, world, port, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import onitama_alphazero_b200 as onb
from onitama_alphazero_b200.sharding import Comm, gather_replay

if rank == 0:
    uid = Comm.unique_id()
    with open(idfile + ".tmp", "wb") as f:
        f.write(uid)
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        assert time.time() - t0 < 60, "no unique id from rank 0"
        time.sleep(0.05)
    uid = open(idfile, "rb").read()

with onb.Context(256, seed=3, device=rank, game_id_base=256 * rank, mcts_max_sims=8) as ctx:
    with Comm(ctx, world, rank, uid) as comm:
        for trial, dst in enumerate((0, world - 1)):
            g = torch.Generator(device="cuda").manual_seed(100 * trial + rank)
            m = [37, 0, 5, 1][(rank + trial) % 4] if trial else 11 + 7 * rank    # different counts per rank, one rank with none
            planes = torch.rand((m, 21, 5, 5), device="cuda", generator=g)
            pi = torch.rand((m, 2, 25), device="cuda", generator=g)
            z = torch.rand((m,), device="cuda", generator=g)
            want = gather_replay(planes, pi, z, dst=dst)
            got = comm.gather_samples(planes, pi, z, dst=dst)
            if rank == dst:
                assert got is not None and all(torch.equal(a, b) for a, b in zip(got, want)), "gather differs from torch.distributed's"
                assert sum(comm.last_counts) == got[0].shape[0]
            else:
                assert got is None and want is None
        # a destination that is too small: EVERY rank gets ONB_E_OVERFLOW and nothing is sent (no rank is left hanging in a send)
        import ctypes as C
        planes = torch.ones((4, 21, 5, 5), device="cuda"); pi = torch.ones((4, 2, 25), device="cuda"); z = torch.ones((4,), device="cuda")
        small = torch.empty((2, 21, 5, 5), device="cuda")
        counts = (C.c_int64 * world)(); total = C.c_int64(0)
        rc = ctx._lib.onb_gather_samples(ctx._h, comm._h, 0, planes.data_ptr(), pi.data_ptr(), z.data_ptr(), 4, small.data_ptr(), small.data_ptr(),
                                         small.data_ptr(), 2, counts, C.byref(total))
        assert rc == -5 and int(total.value) == 4 * world, (rc, total.value)
        # the samples of a native self-play, packed and gathered
        res = ctx.self_play_native(2.0, 8, 256, max_plies=6)
        pk = comm.gather_samples(res["planes"], res["pi"], res["z"], dst=0)
        if rank == 0:
            assert pk[0].shape[0] == sum(comm.last_counts) and comm.last_counts[0] == res["planes"].shape[0]
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
