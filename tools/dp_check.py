"""Data-parallel parity on W GPUs (run under torchrun): a CycleVAEGAN step on a global batch sharded over
the ranks must reproduce the single-GPU step on the full batch: same metrics, and the
all-reduced flat gradient / W equal to the single-GPU gradient.  fp32 parity mode, seeded eps sliced
from a global draw.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from oracle import ref_port as rp
from vcg_b200 import Networks as N
from vcg_b200 import dist as vdist
from vcg_b200 import plan


def make(world_rank=None, world=1):
    torch.manual_seed(1234)
    m = N.CycleVAEGAN(paired=False).cuda()
    m.configure_optimizers(lr=2e-4)
    m.configure_loss(**rp.DEFAULT_LAMBDAS)
    m.train()
    return m


def main():
    rank, local, world = vdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    plan.set_precision(os.environ.get("DP_PREC", "fp32"))
    gb = 2 * world
    batch = rp.synthetic_batch(gb)
    state = {"calls": 0}

    def eps_global(rows):
        def src(shape, device):
            g = torch.Generator().manual_seed(500 + state["calls"])
            state["calls"] += 1
            full = torch.randn(gb, *shape[1:], generator=g)
            return full[rows].to(device)
        return src

    # ---- data parallel
    model = make()
    vdist.attach(model, vdist.GradSync(wire=os.environ.get("DP_WIRE", "fp32")))
    model.optimizer_G.keep_grads = True            # read the reduced gradient after the step
    per = gb // world
    rows = slice(rank * per, (rank + 1) * per)
    N.set_eps_source(eps_global(rows))
    state["calls"] = 0
    steps = int(os.environ.get("DP_STEPS", "1"))
    for _ in range(steps):
        m_dp = model.training_step({"x": batch["x"][rows].cuda(), "y": batch["y"][rows].cuda()})
    g_dp = model.optimizer_G.reduced_grad() * model.optimizer_G.grad_scale
    out = None
    if rank == 0:
        ref = make()
        ref.optimizer_G.keep_grads = True
        N.set_eps_source(eps_global(slice(0, gb)))
        state["calls"] = 0
        for _ in range(steps):
            m_1 = ref.training_step({"x": batch["x"].cuda(), "y": batch["y"].cuda()})
        g_1 = ref.optimizer_G.reduced_grad()
        grabbed, g1 = {"g": g_dp}, {"g": g_1}
        worst = max(abs(m_dp[k] - m_1[k]) / max(abs(m_1[k]), 1e-3) for k in m_1)
        gerr = float((grabbed["g"] - g1["g"]).norm() / g1["g"].norm())
        # weights after the step (generators: deferred Adam; discriminators: immediate)
        # (biases in front of an InstanceNorm have a mathematically zero gradient: Adam turns their ~1e-8 noise into
        #  +-lr-sized steps of random sign, so biases are bounded absolutely by 2*lr instead of relatively)
        werr, berr, wkey = 0.0, 0.0, ""
        for (k, a), (_, b) in zip(model.state_dict().items(), ref.state_dict().items()):
            if not a.dtype.is_floating_point:
                continue
            if a.dim() >= 2:
                e = float((a - b).norm() / (b.norm() + 1e-30))
                if e > werr:
                    werr, wkey = e, k
            else:
                berr = max(berr, float((a - b).abs().max()))
        out = {"world": world, "metrics_worst_rel": worst, "grad_rel_l2": gerr, "weights_worst_rel_l2": werr,
               "weights_worst_key": wkey, "bias_buffer_max_abs_diff": berr,
               "G_loss_dp": m_dp["G_loss"], "G_loss_single": m_1["G_loss"],
               "wire": os.environ.get("DP_WIRE", "fp32"), "steps": steps,
               # bf16 on the wire rounds every rank's gradient to 8 bits of mantissa before the sum: 2^-9 relative per element
               "ok": bool(worst < 1e-4 and gerr < (1e-3 if os.environ.get("DP_WIRE", "fp32") == "fp32" else 6e-3) and
                          werr < (1e-3 if os.environ.get("DP_WIRE", "fp32") == "fp32" else 2e-2) and berr <= 4.1e-4)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
