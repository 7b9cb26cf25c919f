"""Summarise ncu outputs into small text files for profiles/ (run in the authoring container, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv>        -> per-kernel count / total / share
  python tools/ncu_summary.py full <report.ncu-rep>          -> key metrics per profiled launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        n = r["Kernel Name"].split("(")[0].replace("emu::", "").replace("void ", "")[:60]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us(r)
    tot = sum(v[1] for v in agg.values())
    print("# %d launches, %.1f us total (ncu-serialised, cold cache: compare SHARES)" % (len(rows), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s n=%5d total=%10.1f us avg=%8.2f us share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== %s" % name[:100])
        for k in KEYS:
            if k in hdr:
                print("   %-85s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
