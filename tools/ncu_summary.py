#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (text, committed):
   ncu_summary.py launches <launches.csv> <out.txt>          per-kernel launch counts / time shares
   ncu_summary.py report <file.ncu-rep> <out.txt> [regex]    selected metrics per captured launch
   ncu_summary.py traffic <launches.csv> <out.txt> <first_launch> <title>
                                                              per-kernel time + DRAM bytes of the library's launches from
                                                              launch index <first_launch> on (csv with gpu__time_duration.sum,
                                                              dram__bytes_read.sum, dram__bytes_write.sum)
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {sum(v[0] for v in agg.values())} launches, "
                f"{tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:9.1f} {v[1] / tot * 100:6.1f}%\n")


def traffic(path, out, first, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        per.setdefault((row["ID"], row["Kernel Name"]), {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    step = [it for it in list(per.items())[first:] if "veto::" in it[0][1]]

    def short(name):
        m = re.search(r"(\w+_kernel)\s*(<[^(]{0,12})?", name)
        return (m.group(1) + (m.group(2) or "")) if m else re.sub(r"\(.*", "", name)[:48]

    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, name), m in step:
        a = agg[short(name)]
        a[0] += 1
        a[1] += m["gpu__time_duration.sum"] / 1e3
        a[2] += (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / 1e6
    tot, tb = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write(f"# {title}: {len(step)} launches, {tot / 1e3:.2f} ms of kernel time, {tb / 1e3:.1f} GB of DRAM traffic "
                f"(cold-cache, serialised: compare SHARES)\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'dram_MB':>10s} {'TB/s':>6s}\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k:44s} {a[0]:8d} {a[1]:10.1f} {100 * a[1] / tot:6.1f}% {a[2]:10.0f} {a[2] / a[1]:6.2f}\n")


def report(path, out, pattern=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on : {path.split('/')[-1]}\n")
        for n, d in enumerate(data):
            name = d[idx["Kernel Name"]]
            if pattern and not re.search(pattern, name):
                continue
            f.write(f"\n## launch {n}: {name[:150]}\n")
            for m in METRICS:
                if m in idx:
                    f.write(f"{m:100s} {d[idx[m]]:>16s} {units[idx[m]]}\n")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "traffic":
    traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5])
    sys.exit(0)

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
