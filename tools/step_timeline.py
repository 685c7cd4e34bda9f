"""Timeline of one CUDA-graph replayed training step in the DEFAULT schedule (two lanes, weight-gradient side streams,
bucket tails on the optimisers' side streams): every kernel launch of this library bracketed by external CUDA events
inside the captured graph (GraphedStep(profile=True)); start / end offsets are read against the first launch's start
event.  No nsys in the image: this is the per-stream picture it would give.  The event-record nodes themselves cost
time (3 % of a batch-64 step, 14 % of a batch-8 step), so spans are upper bounds; torch library launches (scalar loss
arithmetic, randn) are not bracketed and show up as gaps.

usage: python tools/step_timeline.py [--batch 8] [--out gpurun_out/timeline_b8.csv]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import Networks as N, plan
from vcg_b200.graph import GraphedStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    args = ap.parse_args()
    plan.set_precision("bf16")
    torch.manual_seed(1234)
    m = N.CycleVAEGAN(paired=False).cuda()
    m.configure_optimizers(lr=2e-4)
    m.configure_loss(lambda_kl=1e-5, lambda_gan=1.0, lambda_identity=5.0, lambda_cycle=10.0, lambda_recon=1.0)
    m.train()
    g = torch.Generator().manual_seed(7)
    batch = {"x": torch.rand(args.batch, 3, 256, 256, generator=g).cuda(), "y": torch.rand(args.batch, 3, 256, 256, generator=g).cuda()}
    step = GraphedStep(m, batch, warmup=2, profile=True)
    for _ in range(4):
        step(batch)
    torch.cuda.synchronize()
    t0 = step.records[0][3]
    rows = []
    for kind, work, tag, s, e in step.records:
        rows.append((getattr(s, "vcg_stream", 0), t0.elapsed_time(s) * 1e3, t0.elapsed_time(e) * 1e3, kind, tag))
    names = {}
    for r in rows:
        names.setdefault(r[0], f"s{len(names)}")
    span = max(r[2] for r in rows) - min(r[1] for r in rows)
    with open(args.out, "w") as fh:
        fh.write("stream,start_us,end_us,family,tag\n")
        for r in sorted(rows, key=lambda r: r[1]):
            fh.write(f"{names[r[0]]},{r[1]:.1f},{r[2]:.1f},{r[3]},{r[4]}\n")
    print(f"# batch {args.batch}: {len(rows)} bracketed launches, span {span / 1e3:.3f} ms (first start to last end, with the event nodes)")
    print("# stream  launches  busy_ms  first_start_ms  last_end_ms  idle_inside_ms  largest gaps (us @ ms, after family)")
    for sid, nm in names.items():
        rs = sorted((r for r in rows if r[0] == sid), key=lambda r: r[1])
        busy = sum(r[2] - r[1] for r in rs)
        gaps = [(rs[i + 1][1] - rs[i][2], rs[i][2], rs[i][3]) for i in range(len(rs) - 1)]
        idle = sum(max(0.0, gp[0]) for gp in gaps)
        top = sorted(gaps, reverse=True)[:4]
        print(f"  {nm:5s} {len(rs):6d} {busy / 1e3:9.3f} {rs[0][1] / 1e3:10.3f} {rs[-1][2] / 1e3:12.3f} {idle / 1e3:10.3f}   " +
              "  ".join(f"{gp[0]:.0f}@{gp[1] / 1e3:.2f}:{gp[2]}" for gp in top))
    # how much of the span has 0 / 1 / 2+ streams busy (1 us grid)
    n = int(span) + 2
    occ = [0] * n
    base = min(r[1] for r in rows)
    for r in rows:
        for t in range(int(r[1] - base), min(n, int(r[2] - base) + 1)):
            occ[t] += 1
    for k in range(0, 5):
        c = sum(1 for o in occ if (o == k if k < 4 else o >= 4))
        print(f"# {k if k < 4 else '4+'} launches in flight: {c / 1e3:.3f} ms ({100.0 * c / n:.1f} %)")
    # small-gap statistics: gaps between back-to-back launches of a stream (launch latency of dependent graph nodes)
    allg = []
    for sid in names:
        rs = sorted((r for r in rows if r[0] == sid), key=lambda r: r[1])
        allg += [q[1] - p[2] for p, q in zip(rs[:-1], rs[1:]) if 0 <= q[1] - p[2] < 50]
    allg.sort()
    if allg:
        print(f"# back-to-back gaps < 50 us: {len(allg)}, median {allg[len(allg) // 2]:.1f} us, sum {sum(allg) / 1e3:.3f} ms")


if __name__ == "__main__":
    main()
