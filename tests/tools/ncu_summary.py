"""Summarise an .ncu-rep (ncu --set full) per kernel launch: the counters the north star asks for — issue
utilisation, warp divergence / efficiency, L1 / L2 / HBM traffic — plus the top stall reasons.
  python tests/tools/ncu_summary.py file.ncu-rep [> profiles/xyz_summary.md]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
col = {c: i for i, c in enumerate(h)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}   # -> bytes / microseconds
def g(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    except ValueError:
        return default
M = [
    ("duration us", "gpu__time_duration.sum", 1),
    ("regs/thread", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("lanes/instr", "smsp__thread_inst_executed_per_inst_executed.ratio", 1),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("eligible warps/sched", "smsp__warps_eligible.avg.per_cycle_active", 1),
    ("warp instr (M)", "smsp__inst_executed.sum", 1e-6),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct", 1),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1),
    ("L1 lsu wavefronts % peak", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2 throughput % peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("DRAM read MB", "dram__bytes_read.sum", 1e-6),
    ("DRAM write MB", "dram__bytes_write.sum", 1e-6),
    ("DRAM throughput % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("SM throughput % peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("local load sectors (M)", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", 1e-6),
    ("local store sectors (M)", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", 1e-6),
    ("global load sectors (M)", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 1e-6),
]
stall_cols = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [c for c in h if c.startswith("smsp__average_warp_latency_issue_stalled_") and c.endswith(".ratio")]
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    print("## %s  (launch id %s)" % (name, r[col["ID"]]))
    for label, metric, scale in M:
        v = g(r, metric)
        if v == v:
            print("  %-28s %12.3f" % (label, v * scale))
    st = sorted(((g(r, c, 0.0), c) for c in stall_cols), reverse=True)[:6]
    for v, c in st:
        short = c.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")
        print("  stall %-22s %12.3f" % (short, v))
    print()
