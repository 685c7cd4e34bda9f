"""Summarise .ncu-rep files (ncu --set full) into a small text table for profiles/.
usage: python tools/ncu_summary.py out.txt rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor pipe instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "elapsed cycles"),
    ("smsp__cycles_active.avg", "active cycles (avg SMSP)"),
]


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"== {rep.split('/')[-1]}  ({len(data)} launches profiled; values per launch)"]
    ki = hdr.index("Kernel Name")
    lines.append(f"kernel: {data[0][ki]}")
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            lines.append(f"  {label:42s} {', '.join(r[i] for r in data)} {units[i]}")
    return "\n".join(lines)


if __name__ == "__main__":
    with open(sys.argv[1], "w") as fh:
        for rep in sys.argv[2:]:
            s = summarise(rep)
            print(s)
            fh.write(s + "\n\n")
