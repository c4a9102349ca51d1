"""Per-launch DRAM traffic of the live kernel chain: summarise an
`ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --csv`
launch list (the L2 is NOT flushed between launches, so the bytes are what the step really pulls from / pushes to HBM).
usage: python tools/l2_traffic.py <file.csv>"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, per = None, collections.OrderedDict()
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    per.setdefault((int(d["ID"]), d["Kernel Name"].split("(")[0].replace("void ", "").replace("ptk::", "")[:28]), {})[
        d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
agg = collections.OrderedDict()
print(f"# {sys.argv[1]}: warm-L2 launch list (ncu --cache-control none)")
print(f"{'id':>4s} {'kernel':28s} {'us':>7s} {'rd MB':>8s} {'wr MB':>8s} {'L2 hit %':>8s}")
for (i, k), v in per.items():
    us, rd, wr, hit = v["gpu__time_duration.sum"] / 1e3, v["dram__bytes_read.sum"] / 1e6, v["dram__bytes_write.sum"] / 1e6, v["lts__t_sector_hit_rate.pct"]
    print(f"{i:4d} {k:28s} {us:7.1f} {rd:8.1f} {wr:8.1f} {hit:8.1f}")
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] += rd; a[3] += wr
print("# mean per launch")
for k, a in agg.items():
    print(f"     {k:28s} n={a[0]:3d} {a[1] / a[0]:7.1f} us  rd {a[2] / a[0]:7.1f} MB  wr {a[3] / a[0]:7.1f} MB")
