#!/usr/bin/env python
"""bench.py -- VAE-CycleGAN 256x256 training images/s on N B200 GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W                    (our sm_100a path)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                              (the reference's CPU path)

Workload (BASELINE.json configs[4], the config the metric is quoted on): full VAE-CycleGAN =
CycleVAEGAN(paired=False): 2 VAE generators + 2 discriminators, cycle + KL + LSGAN losses, 256x256,
GLOBAL batch 64 sharded over the N GPUs (strong scaling), synthetic U[0,1) images, random-init weights,
lr 2e-4, Adam(0.5, 0.999), default lambdas.  A step = model.training_step(batch): forward, both backward
sweeps, both Adam updates, metrics read back.

Prints ONE JSON line (rank 0).  `value` times the step with the batch already resident in HBM;
`e2e` times the same public API call with the batch in pinned host memory (every step copies its
100 MB batch host-to-device inside the timed region and reads the metrics back; with CUDA-graph replay the
copy of step i+1 is issued on a side stream during step i, GraphedStep(..., prefetch=), like a pinned-memory
data loader would).  `roofline` is for the dominant kernel family (the tcgen05 implicit
GEMM conv_tc_kernel, forward + data-gradient launches): algorithmic FLOPs (2*M*N*K with the
reference's logical dims) / CUDA-event duration of those launches inside the timed region, against the
measured sustained bf16 peak.  `cpu_baseline` is the oracle port of the reference's training_step on
the host cores (bounded sample)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GLOBAL_BATCH = 64
METRIC = "VAE-CycleGAN 256x256 training images/s (CycleVAEGAN, global batch 64)"
UNIT = "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=2, help="bounded CPU sample: batch of the reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=1, help="1: independent passes of the step on two streams (vcg_b200.lanes)")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph (vcg_b200.graph.GraphedStep)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "MEASURED_PEAKS.json (sustained bf16)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = max([float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()] or [0])
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_step_time(batch, steps, warmup):
    """The reference's training_step (oracle port, same ATen ops) on the host cores."""
    import torch
    from oracle import ref_port as rp
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1234)
    model = rp.RefModel("cyclevaegan", paired=False, lr=2e-4)
    b = rp.synthetic_batch(batch)
    for _ in range(warmup):
        model.training_step(b)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        model.training_step(b)
        ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads()


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ts, threads = cpu_step_time(args.cpu_batch, args.steps, args.warmup)
    total = sum(ts)
    value = args.cpu_batch * len(ts) / total
    sample = (f"CycleVAEGAN(paired=False) training_step, batch {args.cpu_batch} (bounded sample of the global-batch-"
              f"{args.global_batch} workload), fp32, {threads} host threads, oracle port of the reference (a Python "
              "reference cannot travel to the GPU box)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full VAE-CycleGAN (CycleVAEGAN unpaired: 2 VAE generators + 2 discriminators, cycle+KL+LSGAN) "
                               "256x256 training step, global batch %d, per-GPU batch %d" % (args.global_batch, args.global_batch // max(1, args.gpus)),
                   "global_batch": args.global_batch, "parallelism": f"dp{args.gpus}", "latent_dim": 64,
                   "cpu_sample_batch": args.cpu_batch,
                   "note": "reference arm: the reference's CPU path (oracle port, same ATen ops) on a bounded batch of the same workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import vcg_b200  # noqa: F401
    from vcg_b200 import Networks as N
    from vcg_b200 import dist as vdist
    from vcg_b200 import lib, ops, plan

    rank, local, world = vdist.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib.load()
    plan.set_precision(args.precision)
    from vcg_b200 import lanes
    lanes.set_enabled(bool(args.lanes))
    if args.global_batch % world:
        raise SystemExit("global batch must be divisible by the number of GPUs")
    per = args.global_batch // world

    torch.manual_seed(1234)
    model = N.CycleVAEGAN(paired=False).to(dev)
    model.configure_optimizers(lr=2e-4)
    model.configure_loss(lambda_kl=1e-5, lambda_gan=1.0, lambda_identity=5.0, lambda_cycle=10.0, lambda_recon=1.0)
    model.train()
    if world > 1:
        vdist.broadcast_state(model)
        vdist.attach(model)
    g = torch.Generator().manual_seed(7)
    x_all = torch.rand(args.global_batch, 3, 256, 256, generator=g)
    y_all = torch.rand(args.global_batch, 3, 256, 256, generator=g)
    x_host = vdist.shard(x_all, rank, world).contiguous().pin_memory()
    y_host = vdist.shard(y_all, rank, world).contiguous().pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)
    torch.manual_seed(1000 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if args.graph:
        from vcg_b200.graph import GraphedStep
        runner = GraphedStep(model, {"x": x_dev, "y": y_dev}, warmup=2)
    else:
        runner = model.training_step

    def step_resident():
        return runner({"x": x_dev, "y": y_dev})

    def step_e2e():
        # graph mode: GraphedStep copies the pinned host batch straight into its static device buffers
        # and prefetches the next step's batch (here: the same pinned buffers) on a side stream during the replay
        if args.graph:
            hb = {"x": x_host, "y": y_host}
            return runner(hb, prefetch=hb)
        return runner({"x": x_host.to(dev, non_blocking=True), "y": y_host.to(dev, non_blocking=True)})

    for _ in range(args.warmup):
        last = step_resident()
    # ---- timed region (device-resident inputs), conv GEMM launches instrumented with CUDA events
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launch_count()
    if not args.graph:
        ops.prof_begin()
    ms = timed(step_resident, args.steps)
    records = ops.prof_end() if not args.graph else []
    launches = lib.launch_count() - n0
    # ---- end-to-end through the public API with host buffers (right after the resident measurement, same clocks)
    if args.graph:
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
    if args.graph:
        # kernels inside a replayed graph are not re-issued by the library, so they are neither counted nor
        # event-timed there: run the same steps once more eagerly (after the timed region) to count the
        # launches one step consists of and to time the conv GEMM launches with CUDA events
        n0 = lib.launch_count()
        ops.prof_begin()
        ms_eager = timed(lambda: model.training_step({"x": x_dev, "y": y_dev}), args.steps)
        records = ops.prof_end()
        launches = lib.launch_count() - n0
    if not args.graph:
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = args.global_batch * args.steps / (ms / 1e3)
    e2e_value = args.global_batch * args.steps / (ms_e2e / 1e3)
    h2d = 2 * per * 3 * 256 * 256 * 4
    d2h = 4 * len(last)

    # ---- roofline of the dominant kernel family
    fam = {}
    layers = {}
    for kind, flops, tag, s, e in records:
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        dt = s.elapsed_time(e)
        f[0] += flops
        f[1] += dt
        f[2] += 1
        l = layers.setdefault((kind, tag), [0.0, 0.0, 0])
        l[0] += flops
        l[1] += dt
        l[2] += 1
    if rank == 0 and os.environ.get("VCG_BENCH_LAYERS"):
        with open(os.environ["VCG_BENCH_LAYERS"], "w") as fh:
            for (kind, tag), v in sorted(layers.items(), key=lambda kv: -kv[1][1]):
                # conv families: v[0] = FLOPs -> TFLOP/s; xform families: v[0] = bytes -> TB/s (same arithmetic)
                fh.write(f"{kind:16s} {tag:34s} launches/step {v[2] / args.steps:5.1f}  ms/step {v[1] / args.steps:8.3f}  "
                         f"T(FLOP|B)/s {v[0] / max(v[1], 1e-9) / 1e9:8.2f}\n")
    pk = peaks()
    tc_flops = sum(fam.get(k, [0, 0, 0])[0] for k in ("conv_fwd", "conv_dgrad"))
    tc_ms = sum(fam.get(k, [0, 0, 0])[1] for k in ("conv_fwd", "conv_dgrad"))
    tc_n = sum(fam.get(k, [0, 0, 0])[2] for k in ("conv_fwd", "conv_dgrad"))
    achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0
    traffic, traffic_of = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):     # ncu --set full capture of one representative launch (committed evidence)
        tj = json.load(open(tpath))
        traffic, traffic_of = tj.get("conv_tc_kernel_dram_bytes_per_launch"), tj.get("launch")
    roofline = {"kernel": "conv_tc2_kernel / conv_tc_kernel (tcgen05 implicit GEMM, CTA-pair and single-CTA: forward + data-gradient launches)",
                "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_of": traffic_of, "peak_source": pk["source"],
                "launches_per_step": tc_n / max(1, args.steps), "share_of_step": tc_ms / ms,
                "flops_per_launch": tc_flops / max(1, tc_n), "ms_per_launch": tc_ms / max(1, tc_n),
                "families": {k: {"tflops": v[0] / (v[1] / 1e3) / 1e12 if v[1] else 0.0, "ms_per_step": v[1] / args.steps,
                                 "launches_per_step": v[2] / args.steps} for k, v in fam.items()}}
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "full VAE-CycleGAN (CycleVAEGAN unpaired: 2 VAE generators + 2 discriminators, cycle+KL+LSGAN) "
                               "256x256 training step, global batch %d, per-GPU batch %d" % (args.global_batch, per),
                   "global_batch": args.global_batch, "parallelism": f"dp{world}", "latent_dim": 64,
                   "l2": "working set (weights 276 MB bf16 + >10 GB activations per step) exceeds the 126 MB L2; no flush needed",
                   "dead_passes_skipped": True, "cuda_graph": bool(args.graph), "lanes": bool(args.lanes),
                   "roofline_timing": ("conv GEMM launches event-timed in an eager re-run of the same steps right after the "
                                       "graph-replayed timed region" if args.graph else "event-timed inside the timed region")},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "h2d_overlapped_with_previous_step": bool(args.graph)},
        "gpu_launches": launches,
        "clocks": clocks,
        "final_metrics": {k: last[k] for k in ("G_loss", "D_loss", "loss_cycle", "loss_kl")},
    }
    if world == 1 and not args.no_cpu_baseline:
        ts, threads = cpu_step_time(args.cpu_batch, 6, 1)      # ~11 s of CPU work on the box's 16 host threads
        v = args.cpu_batch * len(ts) / sum(ts)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"CycleVAEGAN(paired=False) training_step, batch {args.cpu_batch}, fp32, 1 warm-up + "
                                          f"{len(ts)} timed steps ({sum(ts):.1f} s) of the oracle port on the host cores"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
