"""Worker of test_sharded_failure_does_not_hang (torchrun, one process per GPU): rank CDBG_FAULT_RANK fails inside the first
collective stage; every rank must come back with an error (the others after CDBG_SHARD_TIMEOUT_S) instead of hanging."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bcalm_b200.api import Cdbg, CdbgError  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
c = Cdbg(k=31, abundance_min=2, device=local, rank=rank, world_size=world)
c.shard_init_torch()
ptr, n = c.synth_reads(200_000, 100, 60_000 // world, 10000, 42, first_read=rank * (60_000 // world))
reads = c.copy_to_host(ptr, n)
t0 = time.time()
try:
    c.build(reads)
    outcome = "built"
except CdbgError as e:
    outcome = f"error after {time.time() - t0:.1f} s: {e}"
print(f"FAULT_RESULT rank {rank}: {outcome}", flush=True)
os._exit(0)      # the communicator is in an undefined state: no clean teardown
