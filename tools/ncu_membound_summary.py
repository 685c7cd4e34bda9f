"""HBM evidence for the memory-bound kernels: per-launch DRAM bytes and GB/s from one `ncu --set full` report.
usage: python tools/ncu_membound_summary.py report.ncu-rep|report.csv [out.txt] [--peak 6546.2]
(a .csv is the output of `ncu -i report.ncu-rep --page raw --csv`, made on the GPU box so that only the small text
file has to travel back)

For every profiled launch: kernel, grid, duration, dram__bytes_read.sum + dram__bytes_write.sum, achieved GB/s and the
fraction of the measured HBM copy bandwidth (MEASURED_PEAKS.json: 6546 GB/s).  Durations under ncu are cold-cache and
serialised; the GB/s column is bytes / that duration."""
import csv
import io
import json
import os
import subprocess
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(v, unit):
    u = unit.lower()
    return v * {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3, "second": 1e6}.get(u, 1)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    peak = 6546.2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    if "--peak" in sys.argv:
        peak = float(sys.argv[sys.argv.index("--peak") + 1])
    if args[0].endswith(".csv"):
        out = open(args[0]).read()
    else:
        out = subprocess.run(["ncu", "-i", args[0], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    out = out[out.index('"ID"'):] if '"ID"' in out else out
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "launch__grid_size", "launch__block_size") if k in hdr}
    extra = [k for k in ("sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                         "smsp__issue_active.avg.pct", "launch__registers_per_thread") if k in hdr]
    lines = [f"# {os.path.basename(args[0])}: ncu --set full, {len(data)} launches; HBM peak {peak:.0f} GB/s (measured copy bandwidth)",
             f"# {'kernel':58s} {'grid':>7s} {'us':>9s} {'MB read':>9s} {'MB written':>10s} {'GB/s':>8s} {'of peak':>8s}  " +
             "  ".join(e.split(".")[0].replace("__", ":")[:22] for e in extra)]
    agg = {}
    for r in data:
        name = r[col["Kernel Name"]]
        us = to_us(num(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]])
        rd = to_bytes(num(r[col["dram__bytes_read.sum"]]), units[col["dram__bytes_read.sum"]])
        wr = to_bytes(num(r[col["dram__bytes_write.sum"]]), units[col["dram__bytes_write.sum"]])
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0.0
        lines.append(f"  {name.replace('void ', '').replace('<unnamed>::', '')[:58]:58s} {r[col['launch__grid_size']]:>7s} {us:9.1f} {rd / 1e6:9.2f} {wr / 1e6:10.2f} {gbs:8.0f} {gbs / peak:8.3f}  " +
                     "  ".join(f"{r[hdr.index(e)]:>22s}" for e in extra))
        short = name.replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(short.split("(")[0][:60], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] += rd + wr
    lines.append("# per kernel family (all profiled launches): launches, total us, total MB, GB/s, fraction of peak")
    for k, (n, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = b / (us * 1e-6) / 1e9 if us > 0 else 0.0
        lines.append(f"  {k[:58]:58s} {n:4d} {us:10.1f} {b / 1e6:10.1f} {gbs:8.0f} {gbs / peak:8.3f}")
    text = "\n".join(lines)
    print(text)
    if len(args) > 1:
        with open(args[1], "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
