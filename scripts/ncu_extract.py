"""Condense an `ncu --set full` report into the handful of numbers the roofline discussion uses (run here, no GPU needed):

    python scripts/ncu_extract.py gpurun_out/X.ncu-rep NPOINTS > profiles/X_selected.csv

One line per captured launch: duration, DRAM bytes read / written, bytes per grid point (NPOINTS = points of the level),
achieved DRAM TB/s, active warps, registers, and the stall reasons per issue that matter for these kernels."""
import csv, io, subprocess, sys
rep, npts = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
def val(r, name):
    i = col[name]; return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
stalls = ["long_scoreboard", "barrier", "short_scoreboard", "mio_throttle", "lg_throttle", "wait", "math_pipe_throttle", "not_selected"]
out = csv.writer(sys.stdout)
out.writerow(["kernel", "grid", "block", "duration_us", "dram_read_GB", "dram_write_GB", "dram_bytes_per_point", "dram_TBps", "warps_active_pct", "regs",
              "l2_hit_pct", "l1_hit_pct"] + ["stall_" + s for s in stalls])
for r in data:
    t = val(r, "gpu__time_duration.sum"); rd = val(r, "dram__bytes_read.sum"); wr = val(r, "dram__bytes_write.sum")
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mgb::", "")
    out.writerow([name, r[col["Grid Size"]], r[col["Block Size"]], f"{t * 1e6:.1f}", f"{rd / 1e9:.4f}", f"{wr / 1e9:.4f}", f"{(rd + wr) / npts:.2f}", f"{(rd + wr) / t / 1e12:.3f}",
                  f"{float(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):.1f}", r[col["launch__registers_per_thread"]],
                  f"{float(r[col['lts__t_sector_hit_rate.pct']]):.1f}", f"{float(r[col['l1tex__t_sector_hit_rate.pct']]):.1f}"]
                 + [f"{float(r[col['smsp__average_warps_issue_stalled_' + s + '_per_issue_active.ratio']]):.2f}" for s in stalls])
