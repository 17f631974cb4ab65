"""Turn the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_ncu.py r01            # reads gpurun_out/{launches.csv,nw_prof.ncu-rep,mh_prof.ncu-rep}
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


os.makedirs(OUT, exist_ok=True)
traffic = {}
for name in ("nw_prof", "mh_prof", "nw2_prof", "mh2_prof"):
    rep = os.path.join(G, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, "%s_%s_metrics.csv" % (tag, name)), "w", newline="") as f:
        w = csv.writer(f)
        cols = [c for c in KEEP if c in idx]
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows:
            w.writerow([r[idx[c]] for c in cols])
    # per-launch DRAM traffic of the longest launch of the capture
    best = max(rows, key=lambda r: float(r[idx["gpu__time_duration.sum"]]))
    dram = to_bytes(best[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
        to_bytes(best[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    key = "nw_warp_kernel_dram_bytes_per_launch" if name.startswith("nw") else "mh_match_kernel_dram_bytes_per_launch"
    traffic[key] = dram
    traffic[key + "_kernel"] = best[idx["Kernel Name"]]
    traffic[key + "_grid"] = best[idx["Grid Size"]] if "Grid Size" in idx else None
if traffic:
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (tools/prof_target.py: NW n=700 "
                       "families, MinHash n=32768 -> 536,854,528 pairs); the capture is a reduced workload, so bytes are per THAT launch")
    tp = os.path.join(OUT, "traffic.json")
    merged = json.load(open(tp)) if os.path.exists(tp) else {}
    merged.update(traffic)  # keep entries written by other captures (e.g. the full config-5 launch)
    with open(tp, "w") as f:
        json.dump(merged, f, indent=1)

lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        k = r[4].split("(")[0].replace("void ", "").replace("dyna::<unnamed>::", "")
        t = float(r[-1]) / 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    with open(os.path.join(OUT, "%s_launch_list_summary.md" % tag), "w") as f:
        f.write("# %s: launch list of `python bench.py --steps 1 --warmup 1 --nw-n 2000 --mh-n 20000 --skip-cpu`\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares).\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, c, t, 100 * t / total))
    with open(os.path.join(OUT, "%s_launches.csv" % tag), "w") as f:
        f.write(open(lp).read())
print("profiles written:", sorted(os.listdir(OUT)))
