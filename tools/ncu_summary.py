"""Summarise ncu output.  usage:
  ncu_summary.py launches <launches.csv>          -> per-kernel count / avg us / share (markdown)
  ncu_summary.py full <file.ncu-rep>              -> selected metrics per captured launch (markdown)"""
import csv, io, subprocess, sys, collections

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; iN = hdr.index("Kernel Name"); iV = hdr.index("Metric Value"); iM = hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[iM] != "gpu__time_duration.sum": continue
        agg.setdefault(r[iN], []).append(float(r[iV].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in agg.items():
        print("| `%s` | %d | %.1f | %.3f |" % (k[:70], len(v), sum(v) / len(v), sum(v) / tot))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    iN = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("\n## `%s`\n" % r[iN][:80])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); print("- %s: %s %s" % (k, r[i], units[i]))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
