"""Aggregate pinned host -> device bandwidth with ALL ranks copying at once (what bounds bench.py's e2e at N GPUs):
    python -m torch.distributed.run --nproc-per-node N tools/h2d_probe_multi.py
Every rank copies a 19.3 MB pinned buffer (one uint8 batch of 128 x 3 x 224 x 224) to its GPU 200 times between two barriers;
rank 0 prints the per-rank and the aggregate rate."""
import os, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl' if world > 1 else 'gloo')
n = 128 * 3 * 224 * 224
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(5):
    d.copy_(hs[0], non_blocking=True)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
reps = 200
for i in range(reps):
    d.copy_(hs[i & 1], non_blocking=True)
torch.cuda.synchronize()
sec = torch.tensor([time.perf_counter() - t0], device='cuda')
gathered = [torch.zeros_like(sec) for _ in range(world)]
if world > 1:
    dist.all_gather(gathered, sec)
else:
    gathered = [sec]
if rank == 0:
    rates = [reps * n / float(s) / 1e9 for s in gathered]
    slowest = max(float(s) for s in gathered)
    print('%d ranks: per-rank H2D %s GB/s | aggregate %.1f GB/s (all bytes / slowest rank) = %.0f k img/s of 150 528-byte images'
          % (world, ' '.join('%.1f' % r for r in rates), world * reps * n / slowest / 1e9, world * reps * 128 / slowest / 1e3), flush=True)
dist.destroy_process_group()
