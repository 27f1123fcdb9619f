"""Per-level operator cost with N GPUs (torchrun): relax colour, residual, restrict, prolong per level and the kernels each
level takes.   torchrun --nproc-per-node N scripts/level_costs_multi.py NX NY NZ [nsmall] [kind]   (NX NY NZ = per-GPU block)"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from mgroms_b200 import nhydro as nh, testcases as tc
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
nx, ny, nz = (int(a) for a in sys.argv[1:4])
nsmall = int(sys.argv[4]) if len(sys.argv) > 4 else 8
kind = sys.argv[5] if len(sys.argv) > 5 else "seamount"
nh.reset_params(); nh.set_params(solver_prec=1e-10, nsmall=nsmall)
nh.nhydro_init(nx, ny, nz, grid[0], grid[1], rank, device=local)
if world > 1:
    box = [nh.comm_unique_id() if rank == 0 else None]; dist.broadcast_object_list(box, src=0); nh.comm_init(box[0])
(dx, dy, zeta, h), vert = tc.bench_geometry(kind, nx, ny, nz, grid[0], grid[1], rank)
nh.nhydro_matrices(dx, dy, zeta, h, None, **vert)
u = torch.zeros((nz, ny + 2, nx + 1), dtype=torch.float64).numpy(); v = torch.zeros((nz, ny + 1, nx + 2), dtype=torch.float64).numpy()
w = torch.full((nz + 1, ny + 2, nx + 2), -1.0, dtype=torch.float64).numpy(); w[0] = 0.0
nh.compute_rhs(None, u, v, w); del u, v, w
nh.solve_p(); nh.solve_p()
if rank == 0: print(f"{world} GPUs, {kind} block {nx}x{ny}x{nz}, nsmall={nsmall}, transport {nh.halo_transport()}: solve {nh.last_solve_ms():.2f} ms, {nh.last_solve_iters()} F-cycles")
L = nh.nlevs(); plan = nh.plan_levels(nx, ny, nz, grid[0], grid[1], rank)
def bar():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
def t(f, n=3):
    f(); best = 1e9
    for _ in range(n):
        bar(); t0 = time.perf_counter(); f(); best = min(best, time.perf_counter() - t0)
    return best * 1e6
tot = 0
for lev in range(1, L + 1):
    d = plan[lev - 1]; ns = 20 if d["nx"] * d["ny"] * d["nz"] < 3e7 else 4
    tr = (t(lambda: nh.relax(lev, ns)) - 15) / (2 * ns)
    tres = t(lambda: nh.compute_residual(lev)) - 25
    visits = lev if lev < L else L; sweeps = 5 if lev < L else 40
    cost = visits * sweeps * 2 * tr + (visits * tres if lev < L else 0)
    line = f"lev {lev} {d['nx']}x{d['ny']}x{d['nz']} on {d['npx']}x{d['npy']} gather={d['gather']} [{nh.relax_path(lev)}]: relax colour {tr:8.1f} us residual {tres:8.1f}"
    if lev < L:
        tf = t(lambda: nh.fine2coarse(lev)) - 15; tc2 = t(lambda: nh.coarse2fine(lev)) - 15
        line += f" restrict {tf:7.1f} prolong(exact mode) {tc2:7.1f}"; cost += (1 + visits) * (tf + tc2)
    tot += cost
    if rank == 0: print(line + f" -> {cost/1e3:.2f} ms/cycle")
if rank == 0: print(f"sum {tot/1e3:.2f} ms per F-cycle (API-call estimates)")
nh.nhydro_clean()
if world > 1: dist.barrier(); dist.destroy_process_group()
