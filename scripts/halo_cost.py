"""Cost of one halo exchange (level lev) per call, 2+ GPUs, launched with torchrun."""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from mgroms_b200 import nhydro as nh, testcases as tc
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid = {2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
nx, ny, nz = 512, 512, 64
nh.set_params(solver_prec=1e-10, nsmall=int(sys.argv[1]) if len(sys.argv) > 1 else 8)
nh.nhydro_init(nx, ny, nz, grid[0], grid[1], rank, device=local)
box = [nh.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(box, src=0); nh.comm_init(box[0])
(dx, dy, zeta, h), vert = tc.bench_geometry("seamount", nx, ny, nz, grid[0], grid[1], rank)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
for lev in range(1, nh.nlevs() + 1):
    us = nh.profile_halo(lev, 200)
    if rank == 0: print(f"lev {lev}: {us:.1f} us per fill_halo(p), back to back on the stream")
nh.nhydro_clean(); dist.destroy_process_group()
