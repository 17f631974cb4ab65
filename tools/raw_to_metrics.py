"""profiles/<out>.txt from `ncu -i X.ncu-rep --page raw --csv` files (the .ncu-rep stays on the GPU box: gpurun brings back
at most 64 MiB).   python tools/raw_to_metrics.py profiles/r02b_kernel_metrics.txt gpurun_out/a_raw.csv gpurun_out/b_raw.csv ..."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
out = []
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        out.append("===== %s" % d["Kernel Name"])
        for k in KEYS:
            if k in d:
                out.append("   %-80s %s %s" % (k, d[k], u[k]))
        pre, post = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        for k in sorted(hdr):
            if k.startswith(pre) and k.endswith(post) and "not_issued" not in k and float(d[k]) >= 0.02:
                out.append("   stall %-60s %.3f" % (k[len(pre):-len(post)], float(d[k])))
open(sys.argv[1], "w").write("\n".join(out) + "\n")
print("\n".join(out))
