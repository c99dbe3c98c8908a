"""Summarise ncu outputs into markdown: python scripts/ncu_summary.py launches.csv [report.ncu-rep kernel_regex ...]"""
import collections, csv, io, subprocess, sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names = rows[hdr]
    ki, vi, ui = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[r[ki].split("(")[0].replace("void ", "")].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | mean us | min | max | share of captured GPU time |\n|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {sum(v)/tot:.1%} |")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def report(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    ki = names.index("Kernel Name")
    sel = [r for r in rows[2:] if kern in r[ki]]
    if not sel:
        return
    r = sel[0]
    print(f"\n`{r[ki][:60]}` (first captured launch of {len(sel)}):\n\n| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in names:
            print(f"| {w} | {r[names.index(w)]} | {units[names.index(w)]} |")


if __name__ == "__main__":
    launches(sys.argv[1])
    a = sys.argv[2:]
    for i in range(0, len(a), 2):
        report(a[i], a[i + 1])
