#!/usr/bin/env python
"""Condense ncu reports / launch lists brought back from the GPU box into committed evidence under profiles/ (no GPU needed):

    python tools/ncu_summary.py <session dir under gpurun_out> <tag>

  <dir>/prof_*.ncu-rep   ->  profiles/<tag>_ncu_<name>_raw.csv   (ncu -i ... --page raw --csv, selected metrics)
  <dir>/launches_*.csv   ->  profiles/<tag>_<name>.csv            (copied) + a per-kernel share table printed as markdown
"""
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        res.append((d.get("Kernel Name", "?"), [(k, d[k], u.get(k, "")) for k in KEEP if k in d]))
    return res


def main():
    src, tag = sys.argv[1], sys.argv[2]
    src = src if os.path.isabs(src) else os.path.join(ROOT, "gpurun_out", src)
    out = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(src)):
        p = os.path.join(src, f)
        if f.endswith(".ncu-rep"):
            name = f[:-len(".ncu-rep")].replace("prof_", "")
            ks = raw_metrics(p)
            dst = os.path.join(out, "%s_ncu_%s_raw.csv" % (tag, name))
            with open(dst, "w") as fh:
                w = csv.writer(fh)
                w.writerow(["kernel", "metric", "value", "unit"])
                for kname, ms in ks:
                    for k, v, u in ms:
                        w.writerow([kname[:120], k, v, u])
            print("wrote", os.path.relpath(dst, ROOT), "(%d launches)" % len(ks))
        elif f.startswith("launches_") and f.endswith(".csv"):
            dst = os.path.join(out, "%s_%s" % (tag, f))
            shutil.copyfile(p, dst)
            rows = [r for r in csv.reader(open(p)) if len(r) > 5 and r[0].isdigit()]
            agg = {}
            for r in rows:
                agg.setdefault(r[4].split("(")[0].replace("void ", "")[-48:], []).append(float(r[-1]) / 1000)
            tot = sum(sum(v) for v in agg.values())
            print("\n| kernel (%s) | launches | total us | share | per launch us |\n|---|---|---|---|---|" % f)
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                print("| `%s` | %d | %.1f | %.1f%% | %s |" % (k, len(v), sum(v), 100 * sum(v) / tot, ", ".join("%.0f" % x for x in v[:12])))
            print("total %.1f us" % tot)


if __name__ == "__main__":
    main()
