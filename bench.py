#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MGROMS pressure solver.

Metric (BASELINE.json): ms per multigrid solve of the non-hydrostatic pressure Poisson problem to relative residual
1e-10 (p = 0 start, the reference's solve_p loop, src/mg_solvers.f90:17-101), plus the achieved HBM bandwidth of the
four level-1 kernels (relax colour, residual + norm, fine2coarse, coarse2fine) against the measured roofline.

    python bench.py --gpus N --steps K --warmup W [--config c5|c2|c3|c4] [--nsmall M] [--impl reference]

Configurations (BASELINE.json `configs`, sizes as SURVEY.md section 8):
  c5 (default)  seamount, WEAK scaling, one 1024x2048x128 block per GPU: 8 GPUs = the north-star 4096x4096x128 grid;
                1 GPU = the largest single-GPU configuration
  c2            seamount 512x512x64 per GPU (configs[1] on one GPU; weak-scaled for N > 1)
  c3            random topography 1024x1024x128 GLOBAL, strong scaling on 1/2/4(/8) GPUs
  c4            cuc stand-in 2048x2048x128 GLOBAL on 8 GPUs (strong-scaled down for N < 8), nsmall = 32 so that the coarse
                levels are agglomerated (mg_gather); --nsmall 8 gives the namelist default (no gather)
A "step" is one full solve (all F-cycles until res/bnorm <= 1e-10) with the right-hand side resident in HBM.
`e2e` is the same solve through the reference-facing call nhydro_solve(nx,ny,nz,rmask,u,v,w) with HOST arrays (pinned):
H2D of u,v,w + compute_rhs + solve + correct_uvw + D2H inside the timed region.
For N > 1 the run first checks the live NCCL / peer-memory communicator against the CPU oracle with the SAME
decomposition on two small cases (both halo transports) and aborts on a mismatch (`parity` in the JSON line).
--impl reference times the reference's own Fortran routines, translated statement by statement to C (oracle/_ref, kind
"reference"; the C++ port of the oracle, kind "port", only if that library is missing), one thread per MPI rank over a
thread-emulated MPI, on the same global problem when it fits the host's memory and a time budget; otherwise on the
largest sub-problem that does, flagged `extrapolated`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-10
PROC_GRID = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}     # src/mg_grids.f90:593-594 rank <-> (pi,pj)
BYTES_PER_POINT = dict(relax_colour=68, residual=88, restrict=10, prolong=17)   # SURVEY.md section 8(d), per level-1 point
CONFIGS = {
    "c5": dict(kind="seamount", block=(1024, 2048, 128), scaling="weak", nsmall=8,
               name="BASELINE configs[4]: weak-scaling seamount, 1024x2048x128 per GPU (4096x4096x128 on 8 GPUs)"),
    "c2": dict(kind="seamount", block=(512, 512, 64), scaling="weak", nsmall=8,
               name="BASELINE configs[1]: seamount 512x512x64 per GPU"),
    "c3": dict(kind="rndtopo", glob=(1024, 1024, 128), scaling="strong", nsmall=8,
               name="BASELINE configs[2]: random topography 1024x1024x128 (global), strong scaling"),
    "c4": dict(kind="cuc", glob=(2048, 2048, 128), scaling="strong", nsmall=32,
               name="BASELINE configs[3]: cuc stand-in 2048x2048x128 (global) with coarse-level mg_gather"),
}
ORACLE_BYTES_PER_POINT = 230.0     # host memory of the CPU oracle per level-1 point (all levels, setup arrays, u/v/w), measured
REF_SOLVE_BUDGET_S = 90.0          # reference arm: largest predicted solve time it will really run (set-up takes about as long again)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def local_dims(cfg, n):
    npx, npy = PROC_GRID[n]
    if "block" in cfg:
        return cfg["block"]
    gx, gy, gz = cfg["glob"]
    return gx // npx, gy // npy, gz


def global_dims(cfg, n):
    npx, npy = PROC_GRID[n]
    nx, ny, nz = local_dims(cfg, n)
    return nx * npx, ny * npy, nz


def bind_host_thread_near_gpu(local: int):
    """N > 1: pin this rank's host thread to the CPUs NVML names as closest to its GPU BEFORE the pinned u, v, w buffers are
    allocated, so that they live on that NUMA node (with 4-8 ranks the PCIe legs of nhydro_solve otherwise cross sockets).
    Returns a short description for the JSON line, or None when NVML cannot tell."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return f"{len(cpus)} CPUs near the GPU ({cpus[0]}-{cpus[-1]})"
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places this file executes oracle/): cpu_baseline, --impl reference, and the N>1 parity check
# ------------------------------------------------------------------------------------------------------------------
def cpu_grid(cores: int):
    """Largest power-of-two process grid (npx >= npy) with npx*npy <= cores: the reference runs one MPI rank per core."""
    n = 1
    while n * 2 <= cores:
        n *= 2
    npx = 1
    while npx * npx < n:
        npx *= 2
    return npx, n // npx


def mem_available_bytes():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable"):
                    return float(ln.split()[1]) * 1024.0
    except OSError:
        pass
    return 32e9


def run_oracle(kind, gdims, reps: int, cores: int, nsmall: int):
    """Times the CPU oracle's solve_p on the global grid `gdims`, one thread per emulated rank (native build)."""
    from oracle import pyoracle
    pyoracle.use_native_build(True)
    from oracle.pyoracle import make_case, lib
    nxg, nyg, nz = gdims
    npx, npy = cpu_grid(cores)
    while nxg // npx < 8 or nyg // npy < 8:
        if npx >= npy and npx > 1: npx //= 2
        elif npy > 1: npy //= 2
        else: break
    lib().mgo_set_num_threads(npx * npy)
    t0 = time.perf_counter()
    o = make_case(kind, nxg // npx, nyg // npy, nz, npx, npy, bench_dx=True, solver_prec=TOL, solver_maxiter=50, nsmall=nsmall)
    o.compute_rhs()
    setup_s = time.perf_counter() - t0
    times = []
    for _ in range(reps):
        o.solve_p()
        times.append(o.solve_seconds * 1e3)
    hist = o.history()[0]
    o.close()
    return dict(ms=float(np.mean(times)), iters=len(hist) - 1, npx=npx, npy=npy, final=float(hist[-1]), setup_s=setup_s, reps=reps)


def run_translated_reference(kind, gdims, reps: int, cores: int, nsmall: int):
    """Times solve_p of the reference's OWN Fortran routines -- solver and MPI layer, translated statement by statement to C
    (oracle/f90toc.py; no Fortran compiler in the image), -O3 AVX2+FMA build -- on the global grid `gdims`, one thread per MPI
    rank over the thread-emulated MPI of oracle/ref/ref_mpi.c.  Same interface as run_oracle."""
    from oracle import refdriver
    refdriver.use_fast_build(True)
    from mgroms_b200 import testcases as tc
    nxg, nyg, nz = gdims
    npx, npy = cpu_grid(cores)
    while nxg // npx < 8 or nyg // npy < 8:
        if npx >= npy and npx > 1: npx //= 2
        elif npy > 1: npy //= 2
        else: break
    nx, ny, N = nxg // npx, nyg // npy, npx * npy
    PR = refdriver.RefWorld.PerRank
    t0 = time.perf_counter()
    W = refdriver.RefWorld(nx, ny, nz, npx, npy, solver_prec=TOL, solver_maxiter=50, nsmall=nsmall)
    try:
        geos, vert = [], None
        for r in range(N):
            g, vert = tc.bench_geometry(kind, nx, ny, nz, npx, npy, r)
            geos.append(g)
        W.call("matrices", *[PR([g[q] for g in geos]) for q in range(4)], PR([None] * N), vert["hc"], vert["theta_b"], vert["theta_s"], timeout=3600)
        del geos
        u, v, w = tc.uvw_wminus1(nx, ny, nz)                      # the same (read-only) arrays for every rank: u = v = 0, w = -1
        W.call("compute_rhs", PR([None] * N), u, v, w, timeout=3600)
        del u, v, w
        setup_s = time.perf_counter() - t0
        times, hist = [], None
        for _ in range(reps):
            t1 = time.perf_counter()
            hist = W.call("solve_p", TOL, 50, timeout=7200)[0][0]
            times.append((time.perf_counter() - t1) * 1e3)
    finally:
        W.close()
    return dict(ms=float(np.mean(times)), iters=len(hist) - 1, npx=npx, npy=npy, final=float(hist[-1]), setup_s=setup_s, reps=reps)


PORT_ARM = ("port", "CPU oracle (C++ restatement of the Fortran/MPI path, -O3 -march=native; no Fortran compiler or MPI in this image)")


def cpu_arm():
    """(runner, kind, description): the translated reference when its prebuilt library loads and solves a tiny case, else the C++ port."""
    try:
        from oracle import refdriver
        if refdriver.available():
            run_translated_reference("seamount", (16, 16, 8), 1, 1, 8)          # the library is there and works on this host
            return run_translated_reference, "reference", ("the reference's own Fortran routines (solver + MPI layer) translated statement by statement to C "
                                                          "(oracle/f90toc.py; no Fortran compiler or MPI in this image), gcc -O3 AVX2+FMA, MPI calls emulated between threads")
    except BaseException as e:
        print(f"bench.py: translated reference unavailable ({e!r}); timing the C++ port instead", file=sys.stderr)
    return (run_oracle,) + PORT_ARM


def shrink(gdims, factor):
    """Global grid with `factor` (power of two) times fewer points: the same seamount / topography formula on fewer
    horizontal cells (dx kept), i.e. a bounded sample of the workload."""
    nx, ny, nz = gdims
    f = factor
    while f > 1:
        if ny >= nx: ny //= 2
        else: nx //= 2
        f //= 2
    return nx, ny, nz


def reference_arm(args, rank, cfg):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    g = global_dims(cfg, args.gpus)
    npts = float(g[0]) * g[1] * g[2]
    nsmall = args.nsmall or cfg["nsmall"]
    runner, arm_kind, arm_text = cpu_arm()
    # calibration on a small sample of the same workload: CPU rate in point-cycles per second
    cal_dims = shrink(g, max(1, int(2 ** np.ceil(np.log2(max(1.0, npts / 8.4e6))))))
    cal = runner(cfg["kind"], cal_dims, 1, cores, nsmall)
    cal_pts = float(cal_dims[0]) * cal_dims[1] * cal_dims[2]
    per_pt_ms = cal["ms"] / cal_pts
    avail = mem_available_bytes()
    factor = 1
    while factor < 4096:
        pts = npts / factor
        if pts * ORACLE_BYTES_PER_POINT * 1.15 <= avail and per_pt_ms * pts * 1e-3 * 1.3 <= REF_SOLVE_BUDGET_S:
            break
        factor *= 2
    dims = shrink(g, factor)
    run_pts = float(dims[0]) * dims[1] * dims[2]
    if run_pts <= cal_pts:
        r, dims, run_pts = cal, cal_dims, cal_pts
    else:
        r = runner(cfg["kind"], dims, 1, cores, nsmall)           # one timed solve (tens of seconds of CPU work), whatever --steps says
    scale = npts / run_pts
    value = r["ms"] * scale
    extrap = scale > 1.0 + 1e-9
    sample = (f"{arm_text}, "
              f"{r['npx']}x{r['npy']} MPI ranks = threads on {cores} host cores, {cfg['kind']} {dims[0]}x{dims[1]}x{dims[2]} solved to 1e-10 "
              f"in {r['iters']} F-cycles, 1 timed solve = {r['ms'] / 1e3:.1f} s (setup {r['setup_s']:.0f} s not counted)")
    if extrap:
        sample += (f"; the job's global grid {g[0]}x{g[1]}x{g[2]} does not fit the host's {avail / 1e9:.0f} GB / the {REF_SOLVE_BUDGET_S:.0f} s budget, "
                   f"so the time is EXTRAPOLATED x{scale:.0f} by grid points")
    else:
        sample += f"; this IS the job's global grid {g[0]}x{g[1]}x{g[2]} (nothing extrapolated)"
    conf = config_dict(args, cfg)
    conf["reference_extrapolated"] = bool(extrap); conf["reference_extrapolation_factor"] = float(scale)
    conf["reference_grid"] = f"{dims[0]}x{dims[1]}x{dims[2]}"
    out = {
        "impl": "reference", "metric": "ms per MG solve to rel. residual 1e-10", "value": value, "unit": "ms",
        "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": value, "higher_is_better": False,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": conf,
        "extrapolated": bool(extrap), "fcycles_per_solve": r["iters"],
        "cpu_baseline": {"value": value, "unit": "ms", "cores": r["npx"] * r["npy"], "kind": arm_kind, "sample": sample},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def config_dict(args, cfg):
    n = args.gpus
    npx, npy = PROC_GRID[n]
    nx, ny, nz = local_dims(cfg, n)
    gx, gy, gz = global_dims(cfg, n)
    nsmall = args.nsmall or cfg["nsmall"]
    geo = "dx=dy=156.25 m" if cfg["kind"] != "cuc" else "dx=dy=195 m, hc=250, theta_b=theta_s=6 (synthetic shelf/slope: the reference's NetCDF grid is not shipped)"
    return {"workload": f"{args.config}: {cfg['kind']} {gx}x{gy}x{gz} global = {nx}x{ny}x{nz} per GPU ({cfg['name']}; {geo}), u=v=0, w=-1, "
                        f"F-cycles to res/bnorm<=1e-10, namelist src/nh_namelist with solver_prec=1e-10, nsmall={nsmall} "
                        f"(ns_pre=3, ns_post=2, ns_coarsest=40, RB, linear)",
            "parallelism": f"{npx}x{npy} sub-domains, one per GPU", "nsmall": nsmall,
            "l2": f"inputs larger than L2 (level-1 coefficients {8 * 8 * nx * ny * nz / 1e9:.1f} GB per GPU)"}


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU parity on the live communicator (before the timed region)
# ------------------------------------------------------------------------------------------------------------------
PARITY_CASES = [
    ("seamount", 16, 32, 16, dict(solver_prec=1e-10, solver_maxiter=20, nsmall=16), "random"),   # gathered coarse levels (mg_gather)
    ("rndtopo", 64, 64, 32, dict(solver_prec=1e-10, solver_maxiter=20), "random"),               # level 1 leaves the small / mid kernels
]


def reference_vector(kind, dims, grid, params):
    """Committed output of the translated reference for a decomposed case (tests/golden/ref_golden.json), or None."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_golden.json")) as f:
            cases = json.load(f).get("multi_cases", [])
    except OSError:
        return None
    extras = ",".join(f"{k}={params[k]}" for k in sorted(params) if k not in ("solver_prec", "solver_maxiter"))
    key = f"{kind} {dims[0]}x{dims[1]}x{dims[2]} per rank on {grid[0]}x{grid[1]}" + (f" [{extras}]" if extras else "")
    for c in cases:
        if c["key"] == key:
            return c
    return None


def parity_check(nh, dist, rank, world, local, npx, npy):
    """Same-decomposition comparison with the CPU oracle: cA bit-exact, b bit-exact, residual history
    (|d| <= 1e-10 res + 2e-15), final p (1e-9).  Both halo transports.  Raises on mismatch."""
    import torch
    from oracle.pyoracle import make_case
    from mgroms_b200 import testcases as tc
    worst_h, worst_p, transports, n_ref_vectors = 0.0, 0.0, [], 0
    for transport in ("p2p", "nccl"):
        os.environ["MGB_NO_P2P"] = "0" if transport == "p2p" else "1"
        for kind, nx, ny, nz, prm, uvw in PARITY_CASES:
            o = make_case(kind, nx, ny, nz, npx, npy, uvw=uvw, relax_order=1, **prm)
            o.compute_rhs(); b_or = o.get("b", 1, rank)
            o.solve_p()
            nh.reset_params(); nh.set_params(**prm)
            nh.nhydro_init(nx, ny, nz, npx, npy, rank, device=local)
            box = [nh.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            nh.comm_init(box[0])
            geo = {"seamount": tc.seamount, "rndtopo": tc.rndtopo}[kind](nx, ny, nz, npx, npy, rank)
            vert = tc.SEAMOUNT_VERT
            nh.nhydro_matrices(*geo, None, vert["hc"], vert["theta_b"], vert["theta_s"])
            u, v, w = tc.uvw_random(nx, ny, nz, npx, npy, rank)
            nh.compute_rhs(None, u, v, w)
            ok = all(np.array_equal(nh.get_field("cA", l), o.get("cA", l, rank)) for l in range(1, o.nlevs + 1))
            ok = ok and np.array_equal(nh.get_field("b")[1:-1, 1:-1], b_or[1:-1, 1:-1])
            nh.solve_p()
            hg, ho = nh.history()[0], o.history()[0]
            same_len = len(hg) == len(ho)
            dh = float(np.max(np.abs(hg - ho) / ho)) if same_len else float("inf")
            hist_ok = same_len and bool(np.all(np.abs(hg - ho) <= 1e-10 * ho + 2e-15))
            pg, po = nh.get_field("p"), o.get("p", 1, rank)
            dp = float(np.abs(pg - po).max() / np.abs(po).max())
            used = "p2p" if nh.halo_transport() == "p2p" else "nccl"
            # the same case as produced by the reference's own Fortran text over an emulated MPI (tests/golden/ref_golden.*)
            gv = reference_vector(kind, (nx, ny, nz), (npx, npy), prm)
            if gv is not None:
                import hashlib
                dg = lambda a: hashlib.sha256(np.ascontiguousarray(np.asarray(a, dtype=np.float64) + 0.0).tobytes()).hexdigest()
                ok = ok and all(dg(nh.get_field("cA", l)) == gv["sha_cA"][l - 1][rank] for l in range(1, gv["nlevs"] + 1))
                ok = ok and dg(nh.get_field("b")[1:-1, 1:-1]) == gv["sha_b_interior"][rank]
                if gv["hist_res"][-1] < 1.0:
                    hv = np.array(gv["hist_res"])
                    hist_ok = hist_ok and len(hv) == len(hg) and bool(np.all(np.abs(hg - hv) <= 1e-10 * hv + 2e-15))
                n_ref_vectors += 1
            nh.nhydro_clean(); o.close()
            flags = torch.tensor([0.0 if (ok and hist_ok and dp < 1e-9) else 1.0, dh if np.isfinite(dh) else 1e30, dp], dtype=torch.float64, device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MAX)
            if float(flags[0]) != 0.0:
                raise SystemExit(f"PARITY FAILURE ({transport}/{used}, {kind} {nx}x{ny}x{nz} per rank on {npx}x{npy}): coefficients/rhs bit-exact={ok}, "
                                 f"history ok={hist_ok} (max rel {float(flags[1]):.3e}), p rel {float(flags[2]):.3e}")
            worst_h, worst_p = max(worst_h, float(flags[1])), max(worst_p, float(flags[2]))
            if used not in transports:
                transports.append(used)
    os.environ.pop("MGB_NO_P2P", None)
    return {"checked": True, "max_hist_rel": worst_h, "p_rel": worst_p, "cA_b_bitexact": True, "transport": transports,
            "cases": [f"{k} {nx}x{ny}x{nz} per rank, nsmall={p.get('nsmall', 8)}" for k, nx, ny, nz, p, _ in PARITY_CASES],
            "against": f"CPU oracle with the same {npx}x{npy} decomposition (relax_order=1 form; history tol 1e-10*res+2e-15, p tol 1e-9)"
                       + (f"; and {n_ref_vectors} comparisons with the committed vectors of the translated reference run on {npx}x{npy} emulated MPI ranks "
                          f"(coefficients, rhs bit-exact; history 1e-10)" if n_ref_vectors else ""),
            "reference_vectors_checked": n_ref_vectors}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nsmall", type=int, default=0, help="namelist nsmall (0 = the configuration's default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus not in PROC_GRID:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        reference_arm(args, rank, cfg)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); got {world}")
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from mgroms_b200 import nhydro as nh, testcases as tc

    torch.cuda.set_device(local)
    affinity = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if not os.environ.get("MGB_BENCH_NO_AFFINITY"):
            affinity = bind_host_thread_near_gpu(local)
    npx, npy = PROC_GRID[args.gpus]
    nx, ny, nz = local_dims(cfg, args.gpus)
    kind = cfg["kind"]

    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_check(nh, dist, rank, world, local, npx, npy)

    nh.reset_params()
    nh.set_params(solver_prec=TOL, solver_maxiter=50, nsmall=args.nsmall or cfg["nsmall"])
    nh.nhydro_init(nx, ny, nz, npx, npy, rank, device=local)
    if world > 1:
        box = [nh.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nh.comm_init(box[0])
    (dx, dy, zeta, h), vert = tc.bench_geometry(kind, nx, ny, nz, npx, npy, rank)
    nh.nhydro_matrices(dx, dy, zeta, h, None, vert["hc"], vert["theta_b"], vert["theta_s"])

    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    shapes = ((nz, ny + 2, nx + 1), (nz, ny + 1, nx + 2), (nz + 1, ny + 2, nx + 2))
    u, v, w = (pinned(s) for s in shapes)

    def reset_uvw():                     # u = v = 0, w(:,:,0) = 0, w(:,:,1:nz) = -1  (src/mg_testseamount.f90:119-123)
        u[...] = 0.0; v[...] = 0.0; w[...] = -1.0; w[0] = 0.0
    reset_uvw()
    nh.compute_rhs(None, u, v, w)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident solve: W warm-up + K timed steps
    for _ in range(args.warmup):
        nh.solve_p()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    nh.reset_launch_count()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        nh.solve_p()
        dev_ms += nh.last_solve_ms()       # CUDA events on the solver's stream around the solve_p loop
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = nh.launch_count()
    clocks = sampler.stop()
    iters = nh.last_solve_iters()
    hist = nh.history()[0]
    tt = torch.tensor([dev_ms / args.steps, wall_ms / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step, wall_per_step = float(tt[0]), float(tt[1])

    # ---- end to end through nhydro_solve with host buffers
    e2e = None
    if not args.no_e2e:
        e2e_ms = 0.0
        n_e2e = max(2, min(args.steps, 3))
        for it in range(n_e2e + 1):
            reset_uvw()
            barrier()
            t1 = time.perf_counter()
            nh.nhydro_solve(nx, ny, nz, None, u, v, w)
            torch.cuda.synchronize()
            if it > 0:
                e2e_ms += (time.perf_counter() - t1) * 1e3
        te = torch.tensor([e2e_ms / n_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = (u.nbytes + v.nbytes + w.nbytes) * world
        e2e = {"value": float(te[0]), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d, "solves_timed": n_e2e}

    # ---- roofline of the four level-1 kernels, CUDA events on the solver stream (mgb_profile_level1)
    prof = nh.profile_level1(5 if nx * ny * nz > 1e8 else 10)
    peak, peak_kind = measured_peaks()
    npts = nx * ny * nz
    traffic_db = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_db = json.load(f)
    except Exception:
        pass

    def kernel_line(key):
        ms = prof[key + "_ms"]
        if not ms:
            return None
        ach = BYTES_PER_POINT[key] * npts / (ms * 1e-3) / 1e9
        t = traffic_db.get(key, {})
        per_pt = t.get("dram_bytes_per_point")
        return {"ms": ms, "algorithmic_bytes_per_launch": BYTES_PER_POINT[key] * npts, "achieved": ach, "frac": ach / peak,
                "traffic": (per_pt * npts if per_pt else None), "traffic_source": t.get("source")}
    kernels = {k: kernel_line(k) for k in BYTES_PER_POINT}
    rc = kernels["relax_colour"]
    roofline = {"bound": "hbm", "achieved": rc["achieved"], "peak": peak, "unit": "GB/s", "frac": rc["frac"], "traffic": rc["traffic"],
                "peak_source": peak_kind,
                "kernel": f"relax colour half-sweep at level 1 ({nh.relax_path(1)}; halo and bottom-plane upkeep included), 68 B/pt (SURVEY 8d) x {npts} points",
                "algorithmic_bytes_per_launch": rc["algorithmic_bytes_per_launch"], "kernels": kernels,
                "relax_paths_by_level": [nh.relax_path(l) for l in range(1, nh.nlevs() + 1)]}
    # the whole solve against the HBM bound of SURVEY 8(d): one F-cycle + convergence residual = 910 B per level-1 point at level 1,
    # x 1.3 for the coarser levels = 1.2 kB per level-1 point and cycle
    solve_bytes = 1.2e3 * npts * max(int(iters), 1)
    roofline["whole_solve"] = {"algorithmic_bytes": solve_bytes, "achieved": solve_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
                               "frac": solve_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "note": "1.2 kB per level-1 point and F-cycle (SURVEY 8d) x F-cycles of the solve / measured time per solve"}

    if rank == 0:
        out = {
            "metric": "ms per MG solve to rel. residual 1e-10", "value": ms_per_step, "unit": "ms", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": cfg["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, cfg),
            "fcycles_per_solve": iters, "final_rel_residual": float(hist[-1]), "wall_ms_per_step": wall_per_step,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        }
        if e2e:
            out["e2e"] = e2e
        if affinity:
            out["config"]["host_affinity"] = affinity
        if parity:
            out["parity"] = parity
    nh.nhydro_clean()
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload: the same seamount / topography formula on a grid with fewer horizontal
            # cells (about 10-30 s of CPU work), scaled to the full grid by points
            cores = os.cpu_count() or 1
            factor = max(1, int(2 ** np.ceil(np.log2(max(1.0, npts / 3.4e7)))))
            sd = shrink((nx, ny, nz), factor)
            runner, arm_kind, arm_text = cpu_arm()
            r = runner(kind, sd, 1, cores, args.nsmall or cfg["nsmall"])
            out["cpu_baseline"] = {"value": r["ms"] * factor, "unit": "ms", "cores": r["npx"] * r["npy"], "kind": arm_kind,
                                   "sample": f"{arm_text}, {r['npx']}x{r['npy']} MPI ranks = threads on "
                                             f"{cores} host cores; sample = the same {kind} workload on {sd[0]}x{sd[1]}x{sd[2]} ({r['iters']} F-cycles to 1e-10, "
                                             f"{r['ms']:.0f} ms), scaled x{factor} by grid points to {nx}x{ny}x{nz}"
                                             + ("" if factor > 1 else " (the full grid, nothing scaled)")}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
