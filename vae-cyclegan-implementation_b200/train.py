"""train.py -- the reference's training CLI for the hot path (its train.py:43-128, 584-659).

Keeps the reference's flags and defaults, accepts both spellings of the architecture names (the
parser's `autoencoder, vae, aegan, vaegan, cycleae, cyclevae, cycleaegan, cyclevaegan, doubleae,
doublevae` and the README / north_star's `ae, vae_gan, cycle_vae, vae_cyclegan`), adds `--latent_dim`
(the README uses it, the reference parser lacks it) and a data-parallel launch through torchrun's
environment (RANK / LOCAL_RANK / WORLD_SIZE): `--batch_size` is the GLOBAL batch, sharded over ranks.

Out of scope here: the Hypersim / maps / summer2winter datasets, TensorBoard image
logging and checkpoint management of the reference's main(); the epoch loop below drives
`model.training_step(batch)` exactly like train.py:91-107 on any iterable of {'x','y'} batches, and
ships a synthetic U[0,1) source so the step can be exercised without a dataset."""
from __future__ import annotations

import argparse
import os
import time

import torch

ARCH_ALIASES = {"ae": "autoencoder", "vae_gan": "vaegan", "cycle_vae": "cyclevae", "vae_cyclegan": "cyclevaegan",
                "cycle_ae": "cycleae", "ae_gan": "aegan", "cycle_aegan": "cycleaegan", "cycle_vaegan": "cyclevaegan"}
ARCHS = ["autoencoder", "doubleae", "doublevae", "vae", "aegan", "vaegan", "cycleae", "cyclevae", "cycleaegan",
         "cyclevaegan"]


def canonical_architecture(name):
    name = ARCH_ALIASES.get(name, name)
    if name not in ARCHS:
        raise ValueError(f"Unknown architecture: {name}")
    return name


def create_model(architecture, paired=True, latent_dim=64):
    """arch-name -> class map of the reference (train.py:43-77), plus latent_dim plumbing."""
    from . import Networks as N
    a = canonical_architecture(architecture)
    table = {
        "autoencoder": lambda: N.Autoencoder(), "doubleae": lambda: N.DoubleAutoencoder(),
        "doublevae": lambda: N.DoubleVariationalAutoencoder(latent_dim), "vae": lambda: N.VariationalAutoencoder(latent_dim),
        "aegan": lambda: N.AEGAN(), "vaegan": lambda: N.VAEGAN(latent_dim), "cycleae": lambda: N.CycleAE(paired=paired),
        "cyclevae": lambda: N.CycleVAE(latent_dim, paired=paired), "cycleaegan": lambda: N.CycleAEGAN(paired=paired),
        "cyclevaegan": lambda: N.CycleVAEGAN(latent_dim, paired=paired),
    }
    return table[a]()


def build_parser():
    p = argparse.ArgumentParser(description="Train VAE-CycleGAN models (B200-native hot path)")
    p.add_argument("--architecture", type=str, default="autoencoder", choices=ARCHS + sorted(ARCH_ALIASES))
    p.add_argument("--paired", action="store_true", default=False)
    p.add_argument("--unpaired", dest="paired", action="store_false")
    p.add_argument("--latent_dim", type=int, default=64)
    p.add_argument("--image_size", type=int, default=256)
    p.add_argument("--dataset", type=str, default="synthetic", choices=["synthetic", "hypersim", "summer2winter", "maps"])
    p.add_argument("--data_dir", type=str, default="dataset")
    p.add_argument("--batch_size", type=int, default=5)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--steps_per_epoch", type=int, default=10, help="synthetic source: batches per epoch")
    p.add_argument("--lr", type=float, default=0.0002)
    p.add_argument("--lambda_kl", type=float, default=1e-5)
    p.add_argument("--lambda_gan", type=float, default=1.0)
    p.add_argument("--lambda_identity", type=float, default=5.0)
    p.add_argument("--lambda_cycle", type=float, default=10.0)
    p.add_argument("--lambda_recon", type=float, default=1.0)
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--output_dir", type=str, default="runs")
    p.add_argument("--cuda_graph", action="store_true",
                   help="replay training_step as one CUDA graph (fixed batch shape) and prefetch the next pinned "
                        "host batch on a side stream during the step (vcg_b200.graph.GraphedStep)")
    p.add_argument("--no_cuda", action="store_true", help="accepted for compatibility; there is no CPU path")
    return p


class SyntheticPairs:
    """Iterable of {'x','y'} fp32 NCHW batches in [0,1) (ToTensor range, train.py:189-190), sharded by rank."""

    def __init__(self, global_batch, steps, size, device, rank=0, world=1, same_xy=False, seed=7):
        self.global_batch, self.steps, self.size, self.device = global_batch, steps, size, device
        self.rank, self.world, self.same_xy, self.seed = rank, world, same_xy, seed

    def __len__(self):
        return self.steps

    def __iter__(self):
        from .dist import shard
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.steps):
            x = torch.rand(self.global_batch, 3, self.size, self.size, generator=g)
            y = x if self.same_xy else torch.rand(self.global_batch, 3, self.size, self.size, generator=g)
            yield {"x": shard(x, self.rank, self.world).pin_memory(), "y": shard(y, self.rank, self.world).pin_memory()}


def train_epoch(model, dataloader, device, args=None):
    """The reference's epoch loop around the hot path (train.py:80-128): H2D copy, training_step, metric sums."""
    model.train()
    sums, n = {}, 0
    use_graph = bool(getattr(args, "cuda_graph", False))
    it = iter(dataloader)
    nxt = next(it, None)
    while nxt is not None:
        batch, nxt = nxt, next(it, None)
        if use_graph:
            # one graph per model (fixed batch shape); the look-ahead batch is copied during the replay
            runner = getattr(model, "_vcg_graphed_step", None)
            if runner is None:
                from .graph import GraphedStep
                dev_batch = {"x": batch["x"].to(device), "y": batch["y"].to(device)}
                runner = model._vcg_graphed_step = GraphedStep(model, dev_batch, warmup=2)
            same_shape = nxt is not None and nxt["x"].shape == batch["x"].shape and nxt["x"].is_pinned()
            metrics = runner(batch, prefetch=nxt if same_shape else None)
        else:
            batch = {"x": batch["x"].to(device, non_blocking=True), "y": batch["y"].to(device, non_blocking=True)}
            metrics = model.training_step(batch)
        if "G_loss" not in metrics:
            raise KeyError("training_step must report 'G_loss' (train.py:100-104)")
        for k, v in metrics.items():
            if isinstance(v, (int, float)):
                sums[k] = sums.get(k, 0.0) + float(v)
        n += 1
    return {k: v / max(1, n) for k, v in sums.items()}


def validate(model, dataloader, device):
    model.eval()
    sums, n = {}, 0
    for batch in dataloader:
        batch = {"x": batch["x"].to(device, non_blocking=True), "y": batch["y"].to(device, non_blocking=True)}
        metrics = model.validation_step(batch)
        for k, v in metrics.items():
            if isinstance(v, (int, float)):
                sums[k] = sums.get(k, 0.0) + float(v)
        n += 1
    return {k: v / max(1, n) for k, v in sums.items()}


def main(args):
    from . import dist as vdist
    from . import plan
    if not torch.cuda.is_available():
        raise SystemExit("train.py: a CUDA device is required (B200 / sm_100a); the reference's CPU path is not reimplemented")
    if args.dataset != "synthetic":
        raise SystemExit("train.py: dataset loaders are outside the accelerated hot path; pass your own iterable of "
                         "{'x','y'} batches to train_epoch(), or use --dataset synthetic")
    rank, local, world = vdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    plan.set_precision(args.precision)
    torch.manual_seed(args.seed)
    arch = canonical_architecture(args.architecture)
    model = create_model(arch, paired=args.paired, latent_dim=args.latent_dim).to(device)
    model.configure_optimizers(lr=args.lr)
    model.configure_loss(lambda_kl=args.lambda_kl, lambda_gan=args.lambda_gan, lambda_identity=args.lambda_identity,
                         lambda_cycle=args.lambda_cycle, lambda_recon=args.lambda_recon)
    if world > 1:
        vdist.broadcast_state(model)
        vdist.attach(model)
    torch.manual_seed(args.seed + 1 + rank)
    data = SyntheticPairs(args.batch_size, args.steps_per_epoch, args.image_size, device, rank, world,
                          same_xy=(arch == "autoencoder"))
    for epoch in range(args.epochs):
        t0 = time.time()
        m = train_epoch(model, data, device, args)
        if rank == 0:
            dt = time.time() - t0
            print(f"epoch {epoch + 1}/{args.epochs}  {args.batch_size * len(data) / dt:.1f} img/s  " +
                  "  ".join(f"{k}={v:.5f}" for k, v in sorted(m.items())), flush=True)
    return model


if __name__ == "__main__":
    main(build_parser().parse_args())
