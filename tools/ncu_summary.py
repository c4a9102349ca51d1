"""summaries of ncu exports committed under profiles/:
   ncu_summary.py launches <launches.csv>      per-kernel launch count, mean duration, share, DRAM bytes per launch
   ncu_summary.py full <raw.csv>               key counters of an `ncu -i X.ncu-rep --page raw --csv` export, last launch per kernel"""
import collections, csv, io, sys

def launches(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        k = r["Kernel Name"].split("(")[0].split("<")[0].replace("ptk::", "").replace("void ", "")
        m, u, v = r["Metric Name"], r["Metric Unit"], float(r["Metric Value"].replace(",", ""))
        if m == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1e3)
            cnt[k] += 1
        else:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        agg[k][m] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print("serialised sum of %d launches: %.1f us (every launch alone on the GPU: compare SHARES, not absolutes)" % (sum(cnt.values()), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n, t = cnt[k], a["gpu__time_duration.sum"]
        print(f"{k:30s} n={n:4d} {t / n:8.1f} us/launch  share {t / tot * 100:5.1f} %  dram read {a['dram__bytes_read.sum'] / n / 1e6:7.1f} MB  write {a['dram__bytes_write.sum'] / n / 1e6:7.1f} MB per launch")

def full(path):
    rd = list(csv.reader(open(path)))
    hdr, units = rd[0], rd[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    seen = collections.OrderedDict()
    for r in rd[2:]:
        seen.setdefault(r[idx["Kernel Name"]].split("(")[0].replace("ptk::", "").replace("void ", ""), []).append(r)
    for k, rs in seen.items():
        print("== %s   (%d launches captured, last one shown; grid %s block %s)" % (k, len(rs), rs[-1][idx["Grid Size"]], rs[-1][idx["Block Size"]]))
        for c in want:
            if c in idx:
                print("   %-82s %s %s" % (c, rs[-1][idx[c]], units[idx[c]]))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
