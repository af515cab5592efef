"""`ncu -i X.ncu-rep --page raw --csv` -> a compact per-launch table (committed under profiles/).
usage: python tools/ncu_table.py in.csv out.txt "header comment" """
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time_us", 1.0), ("dram__bytes_read.sum", "dramR_MB", 1.0), ("dram__bytes_write.sum", "dramW_MB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0), ("launch__registers_per_thread", "regs", 1.0),
        ("smsp__inst_executed.sum", "warp_inst", 1.0)]


def main():
    src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# %s\n# one row per profiled launch (cold cache, serialised by ncu: use for traffic / utilisation, not for step time);\n"
                "# GB/s = (dram read + write) / time\n" % note)
        f.write("%-34s %-14s %-12s" % ("kernel", "grid", "block") + "".join(" %10s" % c[1] for c in COLS) + " %9s\n" % "GB/s")
        for r in body:
            name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
            vals = []
            for m, _, _ in COLS:
                try:
                    v = float(r[ix[m]].replace(",", ""))
                    if m == "gpu__time_duration.sum":
                        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ix[m]], 1.0)
                    if m.startswith("dram__bytes"):
                        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[ix[m]], 1.0)
                except Exception:  # noqa: BLE001
                    v = float("nan")
                vals.append(v)
            gbs = (vals[1] + vals[2]) * 1e6 / (vals[0] * 1e-6) / 1e9 if vals[0] > 0 else 0.0
            f.write("%-34s %-14s %-12s" % (name[:34], r[ix["Grid Size"]].replace(" ", ""), r[ix["Block Size"]].replace(" ", "")) +
                    "".join(" %10.3f" % v if v < 1e6 else " %10.3g" % v for v in vals) + " %9.1f\n" % gbs)


if __name__ == "__main__":
    main()
