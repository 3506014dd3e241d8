"""Summarise ncu outputs (launch list CSV, .ncu-rep) into small text files under profiles/."""
import collections
import csv
import subprocess
import sys


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] == "gpu__time_duration.sum":
                k = d["Kernel Name"].split("(")[0].replace("void ", "")
                v = float(d["Metric Value"].replace(",", "")) * (1e-3 if d["Metric Unit"] == "ns" else 1)
                agg[k][0] += 1
                agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)",
           "total_us %.1f" % tot]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-34s launches=%5d  total_us=%10.1f  share=%5.1f%%  avg_us=%.1f" % (k[:34], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def rep(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# `ncu --set full --clock-control none` (%s): one block per captured launch" % path.split("/")[-1]]
    for r in rows[2:]:
        out.append("## " + r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append("  %-80s %s %s" % (w, r[i], units[i]))
    return "\n".join(out)


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    open(dst, "w").write((launch_list(src) if kind == "launches" else rep(src)) + "\n")
