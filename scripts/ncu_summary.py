"""Summarise ncu output brought back in gpurun_out/:
  python scripts/ncu_summary.py launches <launches.csv>            # per-kernel share of the step
  python scripts/ncu_summary.py full <report.ncu-rep> [grid-substr] # key metrics of a --set full capture"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else v * 1e3 if r["Metric Unit"] == "ms" else v
        key = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "") + " grid=" + r["Grid Size"]
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{t / tot * 100:5.1f}%  n={c:4d}  {t / c:9.1f} us/launch  {k[:100]}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__inst_executed.sum"]


def full(path, sub=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    for r in rows[2:]:
        if sub and sub not in r[gi]:
            continue
        print("---", r[ki][:60], "grid", r[gi])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:82s} {r[i][:24]:>16s} {units[i]}")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(*sys.argv[2:])
