"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step's kernels by family.
usage: python tools/launch_summary.py launches.csv [out.txt]
The window is cut to ONE step: from the launch after an optimiser tick (adam_tick_kernel of the discriminators' Adam,
the last kernel family of a step) to the next such tick, when the list contains two of them; otherwise the whole list."""
import collections
import csv
import io
import re
import sys


def load(fn):
    txt = open(fn).read()
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    hdr = rows[0]
    ki, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    out = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[ui], 1.0)
        out.append((r[ki], v))
    return out


def family(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    base = m.group(1) if m else name
    if base.startswith("at::") or base.startswith("cub") or "cublas" in name or base in ("gemmk1_kernel", "dot_kernel", "reduce_1Block_kernel"):
        inner = re.search(r"at::native::(\w+)|(\w+Functor)", name)
        return "torch/" + (inner.group(1) or inner.group(2) if inner else base.split("::")[-1])
    t = m.group(2) or ""
    return base + (t if len(t) < 28 else "")


def main():
    L = load(sys.argv[1])
    ticks = [i for i, (n, _) in enumerate(L) if "adam_tick" in n]
    # a step has two ticks (generators, discriminators); the discriminators' tick is followed by a short tail
    note = f"whole list ({len(L)} launches)"
    if len(ticks) >= 3:
        a, b = ticks[-3], ticks[-1]
        L, note = L[a:b], f"launches {a}..{b} of the list = one training step ({b - a} launches)"
    agg = collections.OrderedDict()
    for n, v in L:
        a = agg.setdefault(family(n), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in L)
    lines = [f"# {sys.argv[1]}: {note}; sum of kernel durations {tot / 1e3:.3f} ms (serialised, cold-cache under ncu)"]
    ours = sum(v for k, (c, v) in agg.items() if not k.startswith("torch/"))
    lines.append(f"# kernels of libvcg_b200.so: {sum(c for k, (c, v) in agg.items() if not k.startswith('torch/'))} launches, "
                 f"{ours / 1e3:.3f} ms; torch library kernels: {sum(c for k, (c, v) in agg.items() if k.startswith('torch/'))} launches, "
                 f"{(tot - ours) / 1e3:.3f} ms")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:64]:64s} launches {c:5d}  ms {v / 1e3:9.3f}  share {100 * v / tot:5.1f} %  avg us {v / c:8.1f}")
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
