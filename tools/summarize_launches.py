"""profiles/<tag>_launch_list_summary.md from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
   python tools/summarize_launches.py gpurun_out/r2r_launches.csv r02 "python bench.py --steps 2 --warmup 1 --skip-cpu" """
import collections
import csv
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        rows.append((r["Kernel Name"], ms))
tot = collections.Counter()
cnt = collections.Counter()
for k, ms in rows:
    k = re.sub(r"\(anonymous namespace\)::|dyna::|<unnamed>::|void ", "", k)
    k = re.sub(r"\(.*$", "", k)
    k = re.sub(r"\(int\)|\(bool\)", "", k)
    tot[k] += ms
    cnt[k] += 1
total = sum(tot.values())
out = ["# %s: launch list of `%s`" % (tag, cmd), "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares). %d launches, %.1f ms in total."
       % (len(rows), total), "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, ms in tot.most_common():
    if ms / total < 0.0005:
        continue
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (k[:150], cnt[k], ms, 100 * ms / total))
with open(os.path.join(ROOT, "profiles", "%s_launch_list_summary.md" % tag), "w") as f:
    f.write("\n".join(out) + "\n")
shutil.copy(src, os.path.join(ROOT, "profiles", "%s_launches.csv" % tag))
print("\n".join(out[:14]))
