"""Summarises ncu reports (read here, without a GPU) into markdown for profiles/.

    python tools/ncu_summary.py <title> <report.ncu-rep> [...]        # --set full captures
    python tools/ncu_summary.py --launches <launches.csv>             # gpu__time_duration launch list -> shares
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def full(title, path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    print(f"\n## {title} ({path.split('/')[-1]}) -- ncu --set full --clock-control none")
    print("kernel | " + " | ".join(f"{hdr[c].split('.')[0].replace('__', ' ')} [{units[c]}]" for c in cols))
    print("---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.replace("cfb::<unnamed>::", "").replace("void ", "")[:70]
        print(name + " | " + " | ".join(f"{float(r[c]):.3f}" if r[c].replace('.', '', 1).replace('e', '', 1).replace('-', '').replace('+', '').isdigit() else r[c] for c in cols))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = r[kn].replace("cfb::<unnamed>::", "").replace("void ", "")
        name = name.split("(")[0][:64]
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
    print(f"\n## launch list of one cfg2 step ({path.split('/')[-1]}): gpu__time_duration.sum, cold-cache and serialised -- compare SHARES")
    print(f"kernel | launches | total [{unit}] | share of step")
    print("---|---|---|---")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f} %")
    print(f"total | {sum(v[0] for v in agg.values())} | {tot:.0f} | 100 %")


def traffic(path, regex, out_json):
    """Average DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels whose name matches
    `regex` in a --set full report -> a small JSON that bench.py copies into roofline.traffic."""
    import json
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for r in rows[2:]:
        if re.search(regex, r[kn]):
            tot += float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            n += 1
    json.dump({"kernel_regex": regex, "launches": n, "traffic_bytes_per_launch": tot / max(n, 1), "report": path.split("/")[-1],
               "how": "ncu --set full --clock-control none (cold L2 per replayed launch), dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches"},
              open(out_json, "w"), indent=1)
    print(f"{n} launches, {tot / max(n, 1) / 1e6:.1f} MB per launch -> {out_json}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "--traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        for p in sys.argv[2:]:
            full(sys.argv[1], p)
