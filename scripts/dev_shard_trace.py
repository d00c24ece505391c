# dev helper (torchrun): sharded C2-per-GPU build with CDBG_TRACE timeline on rank 0
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bcalm_b200.api import Cdbg, MEM_DEVICE
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank != 0: os.environ.pop("CDBG_TRACE", None)
c = Cdbg(k=31, abundance_min=2, device=local, rank=rank, world_size=world); c.shard_init_torch()
L = int(os.environ.get('DEV_L', '100'))
G = int(os.environ.get('DEV_G', str(5_000_000 * world))); n_reads = G * 30 // L // world
ptr, n = c.synth_reads(G, L, n_reads, 10000, 42, first_read=rank * n_reads)
for it in range(4):
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0: print(f"---- iter {it}", file=sys.stderr, flush=True)
    t0 = time.perf_counter(); c.build_raw(ptr, n, MEM_DEVICE, fetch=False); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0: print(f"iter {it}: {dt*1e3:.2f} ms", file=sys.stderr, flush=True)
c.close(); dist.destroy_process_group()
