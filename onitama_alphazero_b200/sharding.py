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


class Comm:
    """onb_comm_*: the library's own NCCL communicator (one per context), so that a host without torch.distributed -- the Rust
    trainer of the reference -- can run the replay gather through the C ABI. `unique_id()` on rank 0, ship the 128 bytes to the other
    ranks over any host channel, then Comm(ctx, n_ranks, rank, id) on every rank."""

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import _lib as L
        buf = (C.c_uint8 * 128)()
        rc = L.load().onb_comm_unique_id(buf)
        if rc != 0:
            raise L.OnbError(rc, "onb_comm_unique_id failed (is NCCL available?)")
        return bytes(buf)

    def __init__(self, ctx, n_ranks, rank, unique_id):
        import ctypes as C
        self.ctx, self.n_ranks, self.rank = ctx, n_ranks, rank
        h = C.c_void_p()
        idb = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        ctx._ck(ctx._lib.onb_comm_create(ctx._h, n_ranks, rank, idb, None, C.byref(h)))
        self._h = h

    def gather_samples(self, planes, pi, z, dst=0):
        """onb_gather_samples: planes [m,21,5,5], pi [m,2,25], z [m] (f32 device tensors on the context's device, m may differ per
        rank) -> the concatenation in rank order on rank dst (None elsewhere). Collective: every rank must call it."""
        import ctypes as C
        import torch
        ctx = self.ctx
        m = int(planes.shape[0])
        dev = planes.device
        with torch.cuda.stream(ctx.torch_stream()):
            planes, pi, z = planes.contiguous(), pi.contiguous(), z.contiguous()
            counts = (C.c_int64 * self.n_ranks)()
            total = C.c_int64(0)
            ctx._ck(ctx._lib.onb_gather_counts(ctx._h, self._h, m, counts, C.byref(total)))   # size the destination's buffers
            t = int(total.value)
            self.last_counts = [int(x) for x in counts]
            if self.rank != dst:
                ctx._ck(ctx._lib.onb_gather_samples(ctx._h, self._h, dst, planes.data_ptr(), pi.data_ptr(), z.data_ptr(), m, None, None, None, 0,
                                                    counts, C.byref(total)))
                ctx.sync()
                return None
            out_p = torch.empty((max(t, 1), 21, 5, 5), dtype=torch.float32, device=dev)
            out_pi = torch.empty((max(t, 1), 2, 25), dtype=torch.float32, device=dev)
            out_z = torch.empty((max(t, 1),), dtype=torch.float32, device=dev)
            ctx._ck(ctx._lib.onb_gather_samples(ctx._h, self._h, dst, planes.data_ptr(), pi.data_ptr(), z.data_ptr(), m, out_p.data_ptr(),
                                                out_pi.data_ptr(), out_z.data_ptr(), max(t, 1), counts, C.byref(total)))
            ctx.sync()
            return out_p[:t], out_pi[:t], out_z[:t]

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.onb_comm_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
