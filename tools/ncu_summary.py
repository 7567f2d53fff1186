"""Condense an .ncu-rep (ncu --set full) into the one-line-per-kernel CSV kept under profiles/,
and optionally write profiles/traffic.json (per-launch DRAM bytes of the fused kernel).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1_x_summary.csv [--traffic POINTS]
"""
import csv, io, json, subprocess, sys
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [hdr.index(m) for m in METRICS if m in hdr]
names = [hdr[c] for c in cols]
ik = hdr.index("Kernel Name")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["Kernel Name"] + names)
    w.writerow([""] + [units[c] for c in cols])
    for r in data:
        w.writerow([r[ik]] + [r[c] for c in cols])
if "--traffic" in sys.argv:
    pts = int(sys.argv[sys.argv.index("--traffic") + 1])
    r = data[-1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
    json.dump({"kernel": r[ik].split("(")[0].strip(), "points_per_launch": pts, "dram_bytes_read": rd,
               "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "dram_bytes_per_point": round((rd + wr) / pts, 3),
               "source": f"{out} (ncu --set full, bench.py default config)"},
              open("profiles/traffic.json", "w"), indent=1)
print(open(out).read())
