#!/usr/bin/env python
"""Turns ncu output brought back from the GPU box into the small tracked files under profiles/.
  python scripts/ncu_summary.py launches gpurun_out/<launches>.csv profiles/<name>.csv      per-kernel launch list -> totals and shares
  python scripts/ncu_summary.py full gpurun_out/<a>.ncu-rep [<b>.ncu-rep ...] profiles/<name>.csv   the --set full metrics DESIGN.md cites"""
import collections, csv, io, subprocess, sys

KEEP = ("gpu__time_duration.sum", "launch__", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fp64_cycles_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput.avg", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_d", "sm__throughput.avg",
        "smsp__warps_eligible.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "derived__smsp__sass_thread_inst_executed_op_dfma")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        d = per.setdefault((name, r[7], r[8]), [])
        d.append(float(r[14]))
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "block", "grid", "launches", "mean_us", "min_us", "max_us", "total_us", "share_of_listed_time"])
        for (name, blk, grid), v in per.items():
            w.writerow([name, blk, grid, len(v), "%.2f" % (sum(v) / len(v) / 1e3), "%.2f" % (min(v) / 1e3), "%.2f" % (max(v) / 1e3), "%.1f" % (sum(v) / 1e3),
                        "%.3f" % (sum(v) / total)])
    print(open(dst).read())


def full(reps, dst):
    cols, table = [], collections.OrderedDict()
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            col = len(cols)
            cols.append(r[hdr.index("Kernel Name")])
            for i, h in enumerate(hdr):
                if h == "Kernel Name" or any(h.startswith(k) for k in KEEP):
                    if "not_issued" in h:
                        continue
                    table.setdefault((h, units[i]), {})[col] = r[i]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(cols))])
        for (h, u), v in table.items():
            w.writerow([h, u] + [v.get(i, "") for i in range(len(cols))])
    print("wrote", dst, len(table), "metrics x", len(cols), "launches")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2:-1], sys.argv[-1])
