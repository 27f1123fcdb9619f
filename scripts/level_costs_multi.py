"""Per-level operator cost with N GPUs (torchrun), nsmall from argv."""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from mgroms_b200 import nhydro as nh, testcases as tc
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid = {2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
nx, ny, nz = 512, 512, 64
nh.set_params(solver_prec=1e-10, nsmall=int(sys.argv[1]) if len(sys.argv) > 1 else 256)
nh.nhydro_init(nx, ny, nz, grid[0], grid[1], rank, device=local)
box = [nh.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(box, src=0); nh.comm_init(box[0])
(dx, dy, zeta, h), vert = tc.bench_geometry("seamount", nx, ny, nz, grid[0], grid[1], rank)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
nh.compute_rhs(None, *tc.uvw_wminus1(nx, ny, nz)); nh.solve_p()
if rank == 0: print("solve ms", nh.last_solve_ms(), "iters", nh.last_solve_iters())
L = nh.nlevs(); plan = nh.plan_levels(nx, ny, nz, grid[0], grid[1], rank)
def t(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e6
tot = 0
for lev in range(1, L + 1):
    d = plan[lev - 1]; ns = 20
    tr = (t(lambda: nh.relax(lev, ns)) - 15) / (2 * ns)
    tres = t(lambda: nh.compute_residual(lev)) - 25
    visits = lev if lev < L else L; sweeps = 5 if lev < L else 40
    cost = visits * sweeps * 2 * tr + (visits * tres if lev < L else 0)
    line = f"lev {lev} {d['nx']}x{d['ny']}x{d['nz']} on {d['npx']}x{d['npy']} gather={d['gather']}: relax colour {tr:7.1f} us residual {tres:7.1f}"
    if lev < L:
        tf = t(lambda: nh.fine2coarse(lev)) - 15; tc2 = t(lambda: nh.coarse2fine(lev)) - 15
        line += f" restrict {tf:6.1f} prolong {tc2:6.1f}"; cost += (1 + visits) * (tf + tc2)
    tot += cost
    if rank == 0: print(line + f" -> {cost/1e3:.2f} ms/cycle")
if rank == 0: print(f"sum {tot/1e3:.2f} ms per F-cycle")
nh.nhydro_clean(); dist.destroy_process_group()
