"""2+ ranks: exercise SlabComm.enable_peer_halo / peer exchange on a small slab (debug + check vs NCCL)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from scpn_fusion_core_b200.slab import SlabComm, SlabLevel
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
nr, rows, h = 1000, 40, 12
L = SlabLevel(world * rows + 1, nr, rank * rows, (rank + 1) * rows + (1 if rank == world - 1 else 0),
              0 if rank == 0 else h, 0 if rank == world - 1 else h, 1.0, 1.0, np.zeros(nr))
comm = SlabComm(rank, world)
x = torch.full((L.rows_loc, nr), float(rank + 1), dtype=torch.float64, device="cuda")
x += torch.arange(L.rows_loc, device="cuda", dtype=torch.float64)[:, None] * 0.001
ref = x.clone()
comm.exchange(ref, L, h)          # NCCL reference
torch.cuda.synchronize()
ok = comm.enable_peer_halo(local, h * nr)
print(f"rank {rank}: peer enabled = {ok}", flush=True)
for it in range(5):
    y = x.clone() + it
    r2 = x.clone() + it
    comm.peer_saved, comm.peer = comm.peer, None
    comm.exchange(r2, L, h)
    comm.peer = comm.peer_saved
    comm.exchange(y, L, h)
    torch.cuda.synchronize()
    print(f"rank {rank} it {it}: equal = {bool(torch.equal(y, r2))}", flush=True)
dist.barrier()
dist.destroy_process_group()
