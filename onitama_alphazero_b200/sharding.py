"""Multi-GPU host logic. Games / trees are independent units: rank r owns a contiguous range of GLOBAL game ids and the
RNG is keyed by the global id, so there is no collective on the hot path (mirrors the reference's thread fan-out,
alphazero-training/src/train.rs:218-245). The only exchange is the optional end-of-iteration gather of replay samples
(SelfPlayData, train.rs:27-33,241-245) to the trainer rank."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """[first, first + count) of global game ids owned by `rank`; ranges are contiguous, disjoint and cover n_total."""
    first = n_total * rank // world
    last = n_total * (rank + 1) // world
    return first, last - first


def gather_replay(planes, pi, z, dst=0, group=None):
    """Gather variable-length replay samples (planes [m,21,5,5] f32, pi [m,2,25] f32, z [m] f32) to rank `dst`.
    Works with NCCL (device tensors, NVLink) and gloo (CPU tensors). Returns the concatenation on dst, None elsewhere.
    2 304 B per sample; ordering = rank order, then local order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return planes, pi, z
    rank = dist.get_rank(group)
    m = torch.tensor([planes.shape[0]], dtype=torch.int64, device=planes.device)
    counts = [torch.zeros_like(m) for _ in range(world)]
    dist.all_gather(counts, m, group=group)
    counts = [int(c.item()) for c in counts]
    mmax = max(counts)
    # pack one record per sample so a single padded gather moves everything
    rec = torch.cat([planes.reshape(-1, 525), pi.reshape(-1, 50), z.reshape(-1, 1)], dim=1)
    pad = torch.zeros((mmax, 576), dtype=rec.dtype, device=rec.device)
    pad[:rec.shape[0]] = rec
    if dist.get_backend(group) == "nccl":
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    allrec = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return allrec[:, :525].reshape(-1, 21, 5, 5), allrec[:, 525:575].reshape(-1, 2, 25), allrec[:, 575]
