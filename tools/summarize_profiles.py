"""Turns the scratch artefacts of tools/gpu_profile_round.sh (gpurun_out/) into the committed summaries under
profiles/ (named per round).  Usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

METRICS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
           "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
           "sm__cycles_active.avg"]


def launches(tag):
    src = os.path.join(G, "launches_%s.csv" % tag)
    if not os.path.exists(src):
        src = os.path.join(G, "launches_r1.csv")
    if not os.path.exists(src):
        return
    lines = [l for l in open(src) if l.startswith('"')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "%s_launches.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/one_step.py\n"
                "# = one position-DDPM step + one feature-DDPM step (each repeats 1000x in the benchmark) + one 32-shape decode "
                "chunk, batch 256\n# (cold-cache, serialised: compare SHARES)  launches=%d total=%.1f us\n" % (n, tot))
        f.write("%-72s %6s %12s %10s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %6d %12.1f %10.2f %6.1f%%\n" % (k[:72], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


def ncu_full(tag):
    rep = os.path.join(G, "prof_gemm_tc_r1.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    with open(os.path.join(P, "%s_gemm_tc_ncu_full.txt" % tag), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on --profile-from-start off  (feature-DDPM step at "
                "batch 256, records SA1.att.v [TMA A], SA1.att.w2+softmax, SA1.mlp.res [TMA A], SA1.mlp.conv1; one cold launch each)\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write("%-66s %-8s %s\n" % (m, units[i], " | ".join(r[i][:44] for r in body)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(P, exist_ok=True)
    launches(tag)
    ncu_full(tag)
    for src, dst in (("prof_lat_auto.txt", "%s_records_latent_step.txt"), ("prof_pos_auto.txt", "%s_records_position_step.txt"),
                     ("timeline_lat.txt", "%s_gemm_timeline_latent.txt"), ("timeline_pos.txt", "%s_gemm_timeline_position.txt"),
                     ("bench_r1.log", "%s_bench.json"), ("bench_r1_reference.log", "%s_bench_reference.json"),
                     ("bench_2gpu.log", "%s_bench_2gpu.json")):
        s = os.path.join(G, src)
        if os.path.exists(s):
            if src.endswith(".log"):
                line = [l for l in open(s) if l.startswith("{")]
                if line:
                    with open(os.path.join(P, dst % tag), "w") as f:
                        json.dump(json.loads(line[-1]), f, indent=1)
            else:
                shutil.copy(s, os.path.join(P, dst % tag))


if __name__ == "__main__":
    main()
