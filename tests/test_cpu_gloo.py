"""world_size-2 test of the host-side multi-rank logic over torch.distributed (gloo, CPU): per-rank input generation,
level planning, neighbour symmetry, the unique-id broadcast plumbing and the max-over-ranks reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from mgroms_b200 import nhydro as nh, testcases as tc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        npx, npy, nx, ny, nz = 2, 1, 16, 32, 8
        nh.reset_params(); nh.set_params(nsmall=16)
        plan = nh.plan_levels(nx, ny, nz, npx, npy, rank)
        plans = [None] * world
        dist.all_gather_object(plans, plan)
        # neighbour tables are mutually consistent at every level (S<->N, E<->W, SW<->NE, SE<->NW)
        opposite = {0: 2, 1: 3, 2: 0, 3: 1, 4: 6, 5: 7, 6: 4, 7: 5}
        for lev in range(len(plan)):
            for r in range(world):
                for q, nb in enumerate(plans[r][lev]["neighb"]):
                    if nb >= 0:
                        assert plans[nb][lev]["neighb"][opposite[q]] == r, (lev, r, q)
        # the coarsest levels are gathered onto one rank each: both ranks list the same family
        last = plans[0][-1]
        assert last["gather"] == 1 or any(l["gather"] for l in plans[0])
        glev = [l["gather"] for l in plans[0]].index(1)
        assert plans[0][glev]["members"] == plans[1][glev]["members"] == [0, 1]
        assert plans[0][glev]["key"] == 0 and plans[1][glev]["key"] == 1
        # unique-id style broadcast (bench.py: rank 0 creates the NCCL id, everyone receives the same 128 bytes)
        box = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        assert box[0] == bytes(range(128))
        # per-rank geometry tiles assemble into the single-rank field
        dx, dy, zeta, h = tc.seamount(nx, ny, nz, npx, npy, rank)
        hs = [None] * world
        dist.all_gather_object(hs, h)
        _, _, _, hglob = tc.seamount(nx * npx, ny * npy, nz, 1, 1, 0)
        for r in range(world):
            pi = r % npx
            assert np.array_equal(hs[r], hglob[pi * nx:pi * nx + nx + 2, :]), r
        # max over ranks of a per-rank time
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == 10.0 + world - 1
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
