"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput... --csv`
log: one row per kernel name (last launch of each), time, DRAM bytes, achieved GB/s, SM busy."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = agg.setdefault((row["ID"], row["Kernel Name"]), {})
        d[row["Metric Name"]] = (float(row["Metric Value"].replace(",", "")), row["Metric Unit"])
    seen = collections.OrderedDict()
    for (_id, k), d in agg.items():
        seen.setdefault(k, []).append(d)
    mb = lambda x: x[0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[x[1]]
    us = lambda x: x[0] * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(x[1], 1e-3)
    print("%-58s %3s %9s %9s %9s %8s %6s" % ("kernel", "n", "time_us", "dramR_MB", "dramW_MB", "GB/s", "SM%"))
    total = 0.0
    for k, v in seen.items():
        d = v[-1]
        t = us(d["gpu__time_duration.sum"])
        total += t
        r, w = mb(d["dram__bytes_read.sum"]), mb(d["dram__bytes_write.sum"])
        sm = d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", (float("nan"), ""))[0]
        name = k.replace("void ", "").replace("unnamed>::", "").split("(")[0][:58]
        print("%-58s %3d %9.1f %9.1f %9.1f %8.0f %6.1f" % (name, len(v), t, r, w, (r + w) / t * 1e3, sm))
    print("sum of the last launch of each kernel: %.1f us" % total)


if __name__ == "__main__":
    main(sys.argv[1])
