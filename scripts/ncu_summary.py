"""Summarise gpurun_out/prof_*.ncu-rep (ncu --set full) and launches_*.csv into profiles/ (markdown + csv)."""
import csv, glob, io, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"), ("sm__inst_executed_pipe_fp64.sum", "fp64 instr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %")]
traffic = {}
out = ["# ncu summaries (%s) — `ncu --set full --clock-control none`, one launch per kernel, scripts/gpu_profile.sh\n" % tag]
for rep in sorted(glob.glob("gpurun_out/prof_*_%s.ncu-rep" % tag)):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(r)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    di = hdr.index("gpu__time_duration.sum")
    def dur(v):
        try:
            return float(v[di].replace(",", ""))
        except Exception:
            return -1.0
    vals = max(rows[2:], key=dur)            # several launches captured: report the longest one
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else rep
    out.append("\n## %s\n\n| metric | value |\n|---|---|" % name.split("(")[0])
    for key, label in want:
        if key in hdr:
            i = hdr.index(key)
            out.append("| %s (`%s`) | %s %s |" % (label, key, vals[i], units[i]))
    def num(key):
        if key not in hdr:
            return None
        i = hdr.index(key)
        try:
            return float(vals[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(units[i], 1)
        except Exception:
            return None
    kn = name.split("(")[0].replace("<unnamed>::", "").strip()
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    if rd is not None and wr is not None:
        traffic[kn] = dict(dram_bytes_per_launch=rd + wr, duration_us=num("gpu__time_duration.sum"), sequences_per_launch=int(os.environ.get("S", "64")),
                           warp_instructions=num("smsp__inst_executed.sum"))
lines = [l for l in open("gpurun_out/launches_%s.csv" % tag) if not l.startswith("==")]
agg = collections.OrderedDict(); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    agg.setdefault(k, [0.0, 0]); agg[k][0] += v; agg[k][1] += 1; tot += v
out.append("\n## launch list — eight consecutive steady-state frames (four publishing), 64 sequences, `--metrics gpu__time_duration.sum`\n")
out.append("total %.1f us over %d launches (cold-cache, serialised: compare SHARES)\n\n| kernel | launches | us/launch | share |\n|---|---|---|---|" % (tot / 1e3, sum(v[1] for v in agg.values())))
for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append("| %s | %d | %.1f | %.3f |" % (k, n, v / 1e3 / n, v / tot))
OUT = os.environ.get("OUT", "profiles")
os.makedirs(OUT, exist_ok=True)
open(OUT + "/%s_ncu_summary.md" % tag, "w").write("\n".join(out) + "\n")
import json, shutil
json.dump(dict(tag=tag, how="ncu --set full --clock-control none, longest of the captured launches of each kernel, %s sequences per launch" % os.environ.get("S", "64"),
               kernels=traffic), open(OUT + "/%s_ncu_traffic.json" % tag, "w"), indent=1)
shutil.copy("gpurun_out/launches_%s.csv" % tag, OUT + "/%s_launches.csv" % tag)
print("\n".join(out[:120]))
