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
data loader would).  `roofline` is for the dominant kernel family (the tcgen05 implicit-GEMM convolution
kernels: forward, data-gradient AND weight-gradient launches, thin 7x7 fold kernels included): algorithmic
FLOPs (2*M*N*K with the reference's logical dims) / CUDA-event duration of those launches, measured INSIDE the
graph-replayed timed region (external events captured into the graph, read after the last timed step), against
the measured sustained bf16 peak; per-family figures are reported next to it.  `cpu_baseline` / `--impl
reference` time the UNMODIFIED reference's training_step (oracle/ref_loader.py: /root/reference or the archive
oracle/build_ref.py made; the oracle port only if neither exists) on the host cores (bounded sample).
`gpu_reference` (informational) times the same unmodified reference modules through stock PyTorch on the same
B200 (fp32 with TF32 as shipped, and bf16 autocast + channels_last), at the largest batch that fits."""
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
WORKLOAD = ("full VAE-CycleGAN (CycleVAEGAN unpaired: 2 VAE generators + 2 discriminators, cycle+KL+LSGAN) "
            "256x256 training step")
# BASELINE.json configs[] (1-based here as in DESIGN.md).  5 is the configuration the metric is quoted on and the default;
# the others run through the same code with --config N (parity-test configurations; informational bench lines).
CONFIGS = {
    5: {"cls": "CycleVAEGAN", "kwargs": {"paired": False}, "batch": GLOBAL_BATCH, "same_xy": False, "metric": METRIC,
        "workload": WORKLOAD, "latent_dim": 64, "step_flops_b64": 57.8e12},
    4: {"cls": "CycleVAE", "kwargs": {"paired": False}, "batch": 32, "same_xy": False, "latent_dim": 64,
        "metric": "Cycle-VAE 256x256 training images/s (CycleVAE unpaired, global batch 32)",
        "workload": "Cycle-VAE (CycleVAE unpaired: 2 VAE generators, cycle+KL) 256x256 training step"},
    3: {"cls": "VAEGAN", "kwargs": {}, "batch": 16, "same_xy": False, "latent_dim": 64,
        "metric": "VAE-GAN 256x256 training images/s (VAEGAN, batch 16)",
        "workload": "VAE-GAN (VAE generator + discriminator, translation-L1 + KL + LSGAN) 256x256 training step"},
    2: {"cls": "VariationalAutoencoder", "kwargs": {"latent_dim": 1024}, "batch": 8, "same_xy": True, "latent_dim": 1024,
        "metric": "VAE 256x256 training images/s (VariationalAutoencoder latent_dim 1024, batch 8)",
        "workload": "VAE (latent_dim 1024, lambda_kl 1e-5, reconstruction-L1 + KL) 256x256 training step"},
    1: {"cls": "Autoencoder", "kwargs": {}, "batch": 4, "same_xy": True, "latent_dim": None,
        "metric": "AE 256x256 training images/s (Autoencoder, batch 4)",
        "workload": "Autoencoder (reconstruction-L1) 256x256 training step"},
}
LOSS_KW = dict(lambda_kl=1e-5, lambda_gan=1.0, lambda_identity=5.0, lambda_cycle=10.0, lambda_recon=1.0)


def optimizers(model):
    """the FusedAdam objects of a composite: (optimizer_G, optimizer_D) or (optimizer,)"""
    return [o for o in (getattr(model, n, None) for n in ("optimizer_G", "optimizer_D", "optimizer")) if o is not None]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (1-based): 5 = full VAE-CycleGAN, global batch 64 (default, the headline)")
    ap.add_argument("--global-batch", type=int, default=0, help="0: the configuration's own batch")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=2, help="bounded CPU sample: batch of the reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=1,
                    help="independent passes of the step on two streams (vcg_b200.lanes): 1 on (library default), 0 off")
    ap.add_argument("--gpu-reference", type=int, default=-1,
                    help="time the unmodified reference through stock PyTorch on this GPU (informational): 1 on, 0 off, "
                         "-1 auto (on for the single-GPU run)")
    ap.add_argument("--overlap", type=int, default=-1,
                    help="gradient-bucket tails (unpack, all-reduce, Adam, re-pack) on the optimisers' side streams behind the "
                         "backward pass: 1 on, 0 off (same launches on the calling stream at step()), -1 library default")
    ap.add_argument("--side-streams", type=int, default=0, help="FusedAdam.side_streams (0: library default)")
    ap.add_argument("--wgrad-side", type=int, default=-1, help="plan.set_wgrad_side (-1: library default)")
    ap.add_argument("--l2-prefetch", type=int, default=-1, help="vcg_set_l2_prefetch (A/B; -1: library default)")
    ap.add_argument("--wire", default="bf16", choices=["bf16", "fp32"], help="gradient all-reduce element type")
    ap.add_argument("--profile", type=int, default=1, help="1: external CUDA events around every launch inside the captured graph")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph (vcg_b200.graph.GraphedStep)")
    args = ap.parse_args()
    args.cfg = CONFIGS[args.config]
    if args.global_batch <= 0:
        args.global_batch = args.cfg["batch"]
    return args


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


def reference_model(device="cpu", cfg=None):
    """-> (model with training_step, kind, origin).  kind 'reference': the UNMODIFIED reference class
    (Networks.CycleVAEGAN, Networks.py:1872-2150); 'port': the oracle restatement of it (same ATen ops)."""
    import torch
    from oracle import ref_loader
    from oracle import ref_port as rp
    cfg = cfg or CONFIGS[5]
    net, origin = ref_loader.load()
    torch.manual_seed(1234)
    if net is not None:
        m = getattr(net, cfg["cls"])(**cfg["kwargs"]).to(device)
        m.configure_optimizers(lr=2e-4)
        m.configure_loss(**LOSS_KW)
        m.train()
        return m, "reference", origin
    if cfg["cls"] != "CycleVAEGAN":
        raise RuntimeError("the reference modules are not available (oracle/_ref archive missing): only the headline "
                           "configuration has an oracle-port fallback")
    return rp.RefModel("cyclevaegan", paired=False, lr=2e-4, device=None if device == "cpu" else device), "port", origin


def synthetic_batch(rp, cfg, b):
    batch = rp.synthetic_batch(b)
    if cfg["same_xy"]:
        batch["y"] = batch["x"]
    return batch


def cpu_step_time(batch, steps, warmup, cfg=None):
    """The reference's own training_step on the host cores (all of them)."""
    import torch
    from oracle import ref_port as rp
    torch.set_num_threads(os.cpu_count())
    cfg = cfg or CONFIGS[5]
    model, kind, origin = reference_model("cpu", cfg)
    b = synthetic_batch(rp, cfg, batch)
    for _ in range(warmup):
        model.training_step(b)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        model.training_step(b)
        ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads(), kind, origin


def gpu_reference(dev, global_batch, cfg=None):
    """Stock PyTorch on the same GPU: the unmodified reference modules, .to(cuda), training_step (train.py:385, 91-97).
    Two modes: as shipped (fp32 tensors, cuDNN TF32 allowed by default) and bf16 autocast + channels_last.
    Largest batch <= global_batch that fits; informational (the reference publishes no GPU number)."""
    import torch
    from oracle import ref_port as rp
    out = {}
    cfg = cfg or CONFIGS[5]
    for mode in ("fp32_tf32_default", "bf16_autocast_channels_last"):
        b = global_batch
        while b >= 1:
            model = None
            try:
                torch.cuda.empty_cache()
                model, kind, origin = reference_model(dev, cfg)
                if mode.startswith("bf16") and kind == "reference":
                    model = model.to(memory_format=torch.channels_last)
                batch = {k: v.to(dev) for k, v in synthetic_batch(rp, cfg, b).items()}

                def step():
                    if mode.startswith("bf16"):
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            return model.training_step(batch)
                    return model.training_step(batch)
                step()
                torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    step()
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                best = min(ts)
                out[mode] = {"value": b / best, "unit": UNIT, "batch": b, "ms_per_step": 1e3 * best, "kind": kind,
                             "origin": origin, "timing": "best of 3 steps after 1 warm-up, host clock around synchronize()"}
                break
            except torch.cuda.OutOfMemoryError:
                b //= 2
            except Exception as e:          # informational block: never fail the bench
                out[mode] = {"error": f"{type(e).__name__}: {e}"[:200], "batch": b}
                break
            finally:
                del model
                torch.cuda.empty_cache()
    return out


def _kw(cfg):
    return ", ".join(f"{k}={v}" for k, v in cfg["kwargs"].items())


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.cfg
    ts, threads, kind, origin = cpu_step_time(args.cpu_batch, args.steps, args.warmup, cfg)
    total = sum(ts)
    value = args.cpu_batch * len(ts) / total
    sample = (f"{cfg['cls']}({_kw(cfg)}) training_step, batch {args.cpu_batch} (bounded sample of the global-batch-"
              f"{args.global_batch} workload), fp32, {threads} host threads, "
              + (f"the unmodified reference modules ({origin})" if kind == "reference" else "oracle port of the reference"))
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s, global batch %d, per-GPU batch %d" % (cfg["workload"], args.global_batch, args.global_batch // max(1, args.gpus)),
                   "global_batch": args.global_batch, "parallelism": f"dp{args.gpus}", "latent_dim": cfg["latent_dim"],
                   "cpu_sample_batch": args.cpu_batch,
                   "note": "reference arm: the reference's own CPU path on a bounded batch of the same workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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
    if args.l2_prefetch >= 0:
        lib.check(lib.load().vcg_set_l2_prefetch(args.l2_prefetch), "vcg_set_l2_prefetch")
    plan.set_precision(args.precision)
    if args.wgrad_side >= 0:
        plan.set_wgrad_side(bool(args.wgrad_side))
    from vcg_b200 import lanes
    if args.global_batch % world:
        raise SystemExit("global batch must be divisible by the number of GPUs")
    per = args.global_batch // world
    # two lanes pay most when layers cannot fill the machine (16x16 maps at a small per-GPU batch: 15.7 -> 13.2 ms at
    # batch 8) and never cost: 81.0 -> 80.3 ms at batch 64, 43.3 -> 42.0 at batch 32 (CUDA-graph replay, one B200)
    use_lanes = bool(args.lanes)
    lanes.set_enabled(use_lanes)

    torch.manual_seed(1234)
    cfg = args.cfg
    model = getattr(N, cfg["cls"])(**cfg["kwargs"]).to(dev)
    model.configure_optimizers(lr=2e-4)
    model.configure_loss(**LOSS_KW)
    model.train()
    opts = optimizers(model)
    for o in opts:
        if args.overlap >= 0:
            o.overlap = bool(args.overlap)
        if args.side_streams > 0:
            o.side_streams = args.side_streams
    sync = None
    if world > 1:
        vdist.broadcast_state(model)
        sync = vdist.attach(model, vdist.GradSync(wire=args.wire))
    g = torch.Generator().manual_seed(7)
    x_all = torch.rand(args.global_batch, 3, 256, 256, generator=g)
    y_all = x_all if cfg["same_xy"] else torch.rand(args.global_batch, 3, 256, 256, generator=g)
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
    # ---- timed region (device-resident inputs); kernel launches bracketed by CUDA events
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launch_count()
    if not args.graph:
        ops.prof_begin()
    ms = timed(step_resident, args.steps)
    records, rec_steps = [], 1
    if args.graph:
        launches = runner.launches_per_step * args.steps
    else:
        records = [(k, w, t, s.elapsed_time(e)) for k, w, t, s, e in ops.prof_end()]
        launches = lib.launch_count() - n0
        rec_steps = args.steps
    # ---- end-to-end through the public API with host buffers (right after the resident measurement, same clocks)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if args.graph and args.profile:
        # roofline leg: the SAME step captured once more with an external CUDA event pair around every kernel launch
        # (event-record nodes inside the graph, re-recorded by each replay) and replayed right after the timed region.
        # The timed graph carries no events: 2600 record nodes cost 3 % of a batch-64 step and 14 % of a batch-8 step.
        # The instrumented capture runs the SERIAL schedule (lanes off): with two lanes a launch's event-bracketed
        # duration is its time on half of the SMs next to whatever the other lane runs, which says nothing about the
        # kernel; at this batch size the two schedules differ by < 1 % in step time.
        lanes.set_enabled(False)
        wside_prev = plan._STATE["wgrad_side"]
        plan.set_wgrad_side(False)
        ov_prev = [(o, o.overlap) for o in opts]
        for o, _ in ov_prev:
            o.overlap = False
        try:
            prof_runner = GraphedStep(model, {"x": x_dev, "y": y_dev}, warmup=2, profile=True)
            for _ in range(3):
                prof_runner({"x": x_dev, "y": y_dev})
            records = prof_runner.kernel_times()
            del prof_runner
        finally:
            lanes.set_enabled(use_lanes)
            plan.set_wgrad_side(wside_prev)
            for o, v in ov_prev:
                o.overlap = v
    clocks = sampler.stop() if rank == 0 else None
    value = args.global_batch * args.steps / (ms / 1e3)
    e2e_value = args.global_batch * args.steps / (ms_e2e / 1e3)
    h2d = 2 * per * 3 * 256 * 256 * 4
    d2h = 4 * len(last)

    # ---- roofline of the dominant kernel family (every tensor-core convolution launch: forward, data gradient,
    #      weight gradient, the 7x7 fold kernels), per family next to it
    fam = {}
    layers = {}
    for kind, work, tag, dt in records:
        f = fam.setdefault(kind, [0.0, 0.0, 0])
        f[0] += work
        f[1] += dt
        f[2] += 1
        l = layers.setdefault((kind, tag), [0.0, 0.0, 0])
        l[0] += work
        l[1] += dt
        l[2] += 1
    if rank == 0 and os.environ.get("VCG_BENCH_LAYERS"):
        with open(os.environ["VCG_BENCH_LAYERS"], "w") as fh:
            for (kind, tag), v in sorted(layers.items(), key=lambda kv: -kv[1][1]):
                # conv families: v[0] = FLOPs -> TFLOP/s; xform families: v[0] = bytes -> TB/s (same arithmetic)
                fh.write(f"{kind:16s} {tag:34s} launches/step {v[2] / rec_steps:5.1f}  ms/step {v[1] / rec_steps:8.3f}  "
                         f"T(FLOP|B)/s {v[0] / max(v[1], 1e-9) / 1e9:8.2f}\n")
    pk = peaks()
    tc_kinds = [k for k in fam if k.startswith("conv_")]
    tc_flops = sum(fam[k][0] for k in tc_kinds)
    tc_ms = sum(fam[k][1] for k in tc_kinds)
    tc_n = sum(fam[k][2] for k in tc_kinds)
    achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0
    traffic, traffic_of = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):     # ncu --set full capture of one representative launch (committed evidence)
        tj = json.load(open(tpath))
        traffic, traffic_of = tj.get("conv_tc_kernel_dram_bytes_per_launch"), tj.get("launch")
    step_ms = ms / args.steps
    roofline = {"kernel": "tcgen05 implicit-GEMM convolution kernels (conv_tc2_kernel / conv_tc_kernel / conv_tc_fold_kernel forward + "
                          "data gradient, wgrad_tc2_kernel / wgrad_tc_kernel / wgrad_fold_kernel weight gradient): every launch of a step",
                "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_of": traffic_of,
                "traffic_source": "static: ncu --set full capture committed under profiles/ (not measured in this run)",
                "peak_source": pk["source"],
                "timing": ("external CUDA events around every launch inside a second, instrumented capture of the same step in "
                           "the fully serial schedule (one stream: lanes, weight-gradient side streams and bucket overlap "
                           "off, so that a launch's bracketed duration is the kernel's own), replayed right after the timed "
                           "region; the timed graph itself carries no events" if args.graph else
                           "CUDA events around every launch of the timed steps" +
                           ("; the two lanes run concurrently on half of the SMs each, so a launch's duration is its time on "
                            "its half of the machine" if use_lanes else "")),
                "launches_per_step": tc_n / max(1, rec_steps), "share_of_step": tc_ms / max(1, rec_steps) / step_ms,
                "flops_per_launch": tc_flops / max(1, tc_n), "ms_per_launch": tc_ms / max(1, tc_n),
                "whole_step_tflops": (cfg["step_flops_b64"] * (args.global_batch / 64.0) / world / (step_ms / 1e3) / 1e12
                                      if cfg.get("step_flops_b64") else None),
                "families": {k: {"tflops" if k.startswith("conv_") else "tbytes_per_s":
                                 v[0] / (v[1] / 1e3) / 1e12 if v[1] else 0.0, "ms_per_step": v[1] / rec_steps,
                                 "launches_per_step": v[2] / rec_steps} for k, v in fam.items()}}
    if rank != 0:
        return
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "%s, global batch %d, per-GPU batch %d" % (cfg["workload"], args.global_batch, per),
                   "baseline_config": args.config,
                   "global_batch": args.global_batch, "parallelism": f"dp{world}", "latent_dim": cfg["latent_dim"],
                   "l2": ("working set (weights 276 MB bf16 + >10 GB activations per step) exceeds the 126 MB L2; no flush needed"
                          if args.config == 5 else "weights + saved activations of a step exceed the 126 MB L2; no flush"),
                   "dead_passes_skipped": True, "cuda_graph": bool(args.graph), "lanes": use_lanes, "bucket_overlap": bool(opts[0].overlap),
                   "gradient_wire": (args.wire if world > 1 else None),
                   "wire_bytes_per_step_per_rank": (sum(o.flat_grad().numel() for o in opts) *
                                                    (2 if args.wire == "bf16" else 4) if world > 1 else 0)},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "h2d_overlapped_with_previous_step": bool(args.graph)},
        "gpu_launches": launches,
        "clocks": clocks,
        "final_metrics": {k: last[k] for k in ("total_loss", "G_loss", "D_loss", "loss_cycle", "loss_kl", "loss_recon") if k in last},
    }
    want_gpu_ref = args.gpu_reference == 1 or (args.gpu_reference < 0 and world == 1)
    if want_gpu_ref:
        # free this arm's graph pool and activations first: the reference keeps ~1.6 GB of fp32 activations per pair
        del runner
        model = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        line["gpu_reference"] = gpu_reference(dev, args.global_batch, cfg)
        best = max((v.get("value", 0.0) for v in line["gpu_reference"].values()), default=0.0)
        if best:
            line["gpu_reference"]["ours_over_best_stock_pytorch"] = value / best
    if world == 1 and not args.no_cpu_baseline:
        ts, threads, kind, origin = cpu_step_time(args.cpu_batch, 6, 1, cfg)      # ~11 s of CPU work on the box's host threads
        v = args.cpu_batch * len(ts) / sum(ts)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{cfg['cls']}({_kw(cfg)}) training_step, batch {args.cpu_batch}, fp32, 1 warm-up + "
                                          f"{len(ts)} timed steps ({sum(ts):.1f} s) of " +
                                          (f"the unmodified reference modules ({origin})" if kind == "reference"
                                           else "the oracle port") + " on the host cores"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        _teardown()         # success path only: after an exception the other ranks may sit in a collective, and a barrier
                            # here would hang until the NCCL timeout -- let the launcher tear the group down instead


def _teardown():
    """Leave NCCL cleanly: every rank drains its device and the group is destroyed before the interpreter exits (rank 0
    builds the JSON line after the other ranks have returned, so the barrier is what keeps them alive until then)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        import gc
        gc.collect()                      # captured graphs that hold NCCL kernels go before the communicator does
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
