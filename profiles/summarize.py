"""Turn gpurun_out/ ncu artefacts into the committed text summaries under profiles/.
usage: python profiles/summarize.py launches <launches.csv> | full <prof.ncu-rep>"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
        k = d["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms summed kernel time")
    print("# (ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} n={v[0]:5d} total_ms={v[1] / 1e3:9.3f} avg_us={v[1] / v[0]:9.1f} share={v[1] / tot:.4f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full capture {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d.get("Kernel Name", "")[:100])
        for w in WANT:
            if w in d:
                print(f"    {w:75s} {d[w]:>16s} {units[hdr.index(w)]}")
        stalls = sorted(((float(d[h].replace(",", "")), h) for h in hdr
                         if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")),
                        reverse=True)[:8]
        for v, h in stalls:
            print(f"    stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.3f} warps/issue")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
