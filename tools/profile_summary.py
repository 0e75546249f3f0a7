"""Turn ncu outputs in gpurun_out/ into the committed summaries under profiles/.

  python tools_profile_summary.py launches gpurun_out/launches.csv profiles/r01_launches_b6.md
  python tools_profile_summary.py full gpurun_out/prof.ncu-rep profiles/r01_expand_full.md
"""
import collections
import csv
import subprocess
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi or not r[0].isdigit():
            continue
        name = r[ki].split("(")[0].replace("stcsp::<unnamed>::", "")
        ns = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        seq.append((r[0], name, r[gi], ns))
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if not k.startswith("void at::"))
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none); cold-cache, serialised: compare shares\n\n")
        f.write("| kernel | launches | total us | share of all | share of library kernels |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lib = "" if k.startswith("void at::") else "%.3f" % (v[1] / mine)
            f.write("| %s | %d | %.1f | %.3f | %s |\n" % (k[:70], v[0], v[1] / 1e3, v[1] / tot, lib))
        f.write("\nFirst 60 launches (id, kernel, grid, ns):\n\n```\n")
        for s in seq[:60]:
            f.write("%s %s %s %.0f\n" % s)
        f.write("```\n")


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]


def full(src, dst):
    out = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, %s\n\n" % src)
        for r in rows[2:]:
            f.write("## launch id %s\n\n" % r[0])
            for w in WANT:
                if w in h:
                    f.write("- %s = %s %s\n" % (w, r[h.index(w)], rows[1][h.index(w)]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
