"""Per-source-line instruction / stall attribution from an .ncu-rep (needs -lineinfo builds).
usage: ncu_lines.py report.ncu-rep <kernel-regex> [top]"""
import csv, io, re, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header rows; sections per kernel & file
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ""])
hdr = None
cur_file = ""
kernel_seen = 0
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        continue
    if r[0] in ("File Name", "File Path"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] in ("Line No", "Address", "#"):
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if "Line No" in d and d.get("Line No", "").isdigit():
        key = (cur_file, int(d["Line No"]))
        def f(k):
            try:
                return float(d.get(k, "0").replace(",", "") or 0)
            except ValueError:
                return 0.0
        agg[key][0] += f("Instructions Executed")
        agg[key][1] += f("Thread Instructions Executed")
        agg[key][2] += f("Warp Stall Sampling (All Samples)")
        agg[key][3] = d.get("Source", "")[:90]
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[2] for v in agg.values()) or 1
print("total warp-instr %.3g  stall samples %d" % (ti, ts))
print("by instructions:")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% stall lanes %4.1f  %s:%d  %s" % (100 * v[0] / ti, 100 * v[2] / ts, v[1] / max(v[0], 1), f, l, v[3]))
