"""Multi-GPU plumbing: one process per GPU (torchrun), batch-split data parallelism.

Every operator on the hot path is per-image (BatchNorm is pre-folded, planer/io.py:76-91), so the batch index
-- the outermost factor of the GEMM N dimension in planer/util.py:33,42 -- shards with NO collective on the
forward path.  The only exchange is at load time: rank 0 reads the ``.npy`` weight blob and broadcasts it
once (NCCL over NVLink on GPUs; gloo in the CPU tests).  ``torch.distributed`` is plumbing here, nothing more.
"""
import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def initialized():
    try:
        d = _dist()
        return d.is_available() and d.is_initialized()
    except Exception:
        return False


def rank_world():
    if initialized():
        d = _dist()
        return d.get_rank(), d.get_world_size()
    return 0, 1


def shard_batch(n, rank=None, world=None):
    """Contiguous split of ``n`` images: rank r of W gets [start, stop).  Remainders go to the low ranks."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_host_blob(host, total, src=0):
    """Host-side broadcast of the uint8 weight blob (any backend; used with gloo in the CPU tests)."""
    import torch
    rank, world = rank_world()
    if world == 1:
        return np.ascontiguousarray(host).reshape(-1).view(np.uint8)[:total]
    t = torch.empty(total, dtype=torch.uint8)
    if rank == src:
        t.copy_(torch.from_numpy(np.ascontiguousarray(host).reshape(-1).view(np.uint8)[:total].copy()))
    _dist().broadcast(t, src=src)
    return t.numpy()


def upload_blob(host, total, broadcast=None, src=0):
    """uint8 blob -> device.  With torch.distributed initialised (and ``broadcast`` not False) only ``src``
    needs ``host``; the other ranks receive the bytes through ONE NCCL broadcast."""
    from . import backend as B
    import torch
    rank, world = rank_world()
    do_bcast = world > 1 and broadcast is not False
    if not do_bcast:
        if host is None:
            raise ValueError('load_weights: no weight blob given and no process group to receive it from')
        return B.asarray(np.ascontiguousarray(host).reshape(-1).view(np.uint8)[:total])
    blob = B.empty((total,), np.uint8)
    if rank == src:
        if host is None:
            raise ValueError('load_weights: rank %d is the broadcast source but has no blob' % src)
        blob = B.asarray(np.ascontiguousarray(host).reshape(-1).view(np.uint8)[:total])
    return broadcast_device_blob(blob, src)


def broadcast_device_blob(blob, src=0):
    from . import backend as B
    import torch
    if not initialized() or rank_world()[1] == 1:
        return blob
    t = blob._typed()
    B.synchronize()                                  # the upload ran on the library stream
    _dist().broadcast(t, src=src)
    torch.cuda.current_stream().synchronize()
    return blob


def max_over_ranks(value):
    """MAX all-reduce of a python float (bench timing: the slowest rank defines the step)."""
    if not initialized() or rank_world()[1] == 1:
        return float(value)
    import torch
    d = _dist()
    dev = 'cuda' if d.get_backend() == 'nccl' else 'cpu'
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if initialized() and rank_world()[1] > 1:
        _dist().barrier()


def gather_ints(value):
    """All-gather of one python int per rank (load-time checks: every rank reports the CRC of the weight blob it
    received and of its logits on a shared input; rank 0 compares).  Returns the list indexed by rank."""
    if not initialized() or rank_world()[1] == 1:
        return [int(value)]
    import torch
    d = _dist()
    dev = 'cuda' if d.get_backend() == 'nccl' else 'cpu'
    mine = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(d.get_world_size())]
    d.all_gather(out, mine)
    return [int(t.item()) for t in out]
