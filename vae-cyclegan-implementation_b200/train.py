"""train.py -- the reference's training CLI around the accelerated step (its train.py:43-171, 360-659).

Every flag of the reference parser is accepted with the reference's default (train.py:588-656); the architecture
names also take the README / north_star spellings (`ae, vae_gan, cycle_vae, vae_cyclegan`); additions:
`--latent_dim` (README uses it, the reference parser lacks it), `--dataset synthetic` (seeded U[0,1) tensors, no files),
`--precision`, `--seed`, `--cuda_graph`, `--steps_per_epoch`, and a data-parallel launch through torchrun's
environment (RANK / LOCAL_RANK / WORLD_SIZE): `--batch_size` is the GLOBAL batch, sharded over the ranks.

main() keeps the reference's run layout -- <output_dir>/<arch>_<MMDD_HHMM>_<src>_to_<tgt>_<dataset>/{args.json,
best_model.pth, checkpoint_epoch_N.pth, tensorboard/} -- its resume semantics (continue in the checkpoint's
directory at epoch+1), pretrain transfer flags, initial validation, best-model and periodic saves (rank 0 writes).

train_epoch() / validate() return what the reference's return (train.py:80-171).  The reference runs one more
no-grad forward of the whole model in train() mode after EVERY batch and keeps only the last one for display
(train.py:112-117); that forward consumes bottleneck noise (6 draws for CycleVAEGAN) and runs the spectral-norm
power iteration.  Here the noise of the skipped forwards is drawn and dropped, so the random stream -- and with it
every later step -- is the reference's, and only the last batch's forward is actually computed (the power iteration
is idempotent on unchanged weights, see Networks.SpectralConvParams)."""
from __future__ import annotations

import argparse
import json
import os
import time
from datetime import datetime
from pathlib import Path

import torch

ARCH_ALIASES = {"ae": "autoencoder", "vae_gan": "vaegan", "cycle_vae": "cyclevae", "vae_cyclegan": "cyclevaegan",
                "cycle_ae": "cycleae", "ae_gan": "aegan", "cycle_aegan": "cycleaegan", "cycle_vaegan": "cyclevaegan"}
ARCHS = ["autoencoder", "doubleae", "doublevae", "vae", "aegan", "vaegan", "cycleae", "cyclevae", "cycleaegan",
         "cyclevaegan"]
DATASET_MODALITIES = {"hypersim": ("depth", "normal"), "summer2winter": ("summer", "winter"), "maps": ("satellite", "map"),
                      "synthetic": ("x", "y")}


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
    p.add_argument("--pretrained_doubleae", type=str, default=None)
    p.add_argument("--pretrained_doublevae", type=str, default=None)
    p.add_argument("--data_dir", type=str, default="dataset")
    p.add_argument("--source_modality", type=str, default=None)
    p.add_argument("--target_modality", type=str, default=None)
    p.add_argument("--image_size", type=int, default=256)
    p.add_argument("--test_split", type=float, default=0.1)
    p.add_argument("--val_split", dest="test_split", type=float, help="README spelling of --test_split")
    p.add_argument("--dataset", type=str, default="hypersim", choices=["hypersim", "summer2winter", "maps", "synthetic"])
    p.add_argument("--batch_size", type=int, default=5)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--lr", type=float, default=0.0002)
    p.add_argument("--lambda_kl", type=float, default=1e-5)
    p.add_argument("--lambda_gan", type=float, default=1.0)
    p.add_argument("--lambda_identity", type=float, default=5.0)
    p.add_argument("--lambda_cycle", type=float, default=10.0)
    p.add_argument("--lambda_recon", type=float, default=1.0)
    p.add_argument("--output_dir", type=str, default="runs")
    p.add_argument("--save_freq", type=int, default=10)
    p.add_argument("--log_image_freq", type=int, default=5)
    p.add_argument("--resume", type=str, default=None)
    p.add_argument("--num_workers", type=int, default=1)
    p.add_argument("--no_cuda", action="store_true", help="accepted for compatibility; there is no CPU path")
    # additions
    p.add_argument("--latent_dim", type=int, default=64)
    p.add_argument("--steps_per_epoch", type=int, default=10, help="synthetic source: batches per epoch")
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--seed", type=int, default=1234)
    p.add_argument("--cuda_graph", action="store_true",
                   help="replay training_step as one CUDA graph (fixed batch shape) and prefetch the next pinned "
                        "host batch on a side stream during the step (vcg_b200.graph.GraphedStep)")
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
        pin = torch.cuda.is_available()
        for _ in range(self.steps):
            x = torch.rand(self.global_batch, 3, self.size, self.size, generator=g)
            y = x if self.same_xy else torch.rand(self.global_batch, 3, self.size, self.size, generator=g)
            xs, ys = shard(x, self.rank, self.world), shard(y, self.rank, self.world)
            yield {"x": xs.pin_memory() if pin else xs, "y": ys.pin_memory() if pin else ys}


def _accumulate(sums, metrics):
    for k, v in metrics.items():
        if isinstance(v, (int, float)):
            sums[k] = sums.get(k, 0.0) + float(v)


def _display_forward(model, x, y):
    """the reference's per-batch visualisation forward (train.py:112-117): first output of forward(), no grad"""
    from . import Networks as N
    with torch.no_grad():
        out = model(x) if isinstance(model, (N.Autoencoder, N.VariationalAutoencoder)) else model(x, y)
    return out[0] if isinstance(out, (tuple, list)) else out


def train_epoch(model, dataloader, device, args=None, writer=None, epoch=None):
    """-> (avg G_loss, avg metrics, last display output, last x, last y), train.py:80-128."""
    model.train()
    sums, n, total = {}, 0, 0.0
    use_graph = bool(getattr(args, "cuda_graph", False))
    last_x = last_y = last_out = None
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
            same_shape = nxt is not None and nxt["x"].shape == batch["x"].shape and not nxt["x"].is_cuda and nxt["x"].is_pinned()
            metrics = runner(batch, prefetch=nxt if same_shape else None)
            x, y = runner.x, runner.y
        else:
            x, y = batch["x"].to(device, non_blocking=True), batch["y"].to(device, non_blocking=True)
            metrics = model.training_step({"x": x, "y": y})
        if "G_loss" not in metrics:
            raise KeyError("training_step must report 'G_loss' (train.py:100-104)")
        total += metrics["G_loss"]
        _accumulate(sums, metrics)
        n += 1
        last_x, last_y = x, y
        if nxt is None:
            last_out = _display_forward(model, x, y)
        else:
            skip = getattr(model, "skip_forward_noise", None)
            if skip is not None:
                skip(x, y)              # same RNG consumption as the forward the reference runs here
    if n == 0:
        nan = float("nan")
        return nan, {k: nan for k in sums}, None, None, None
    return total / n, {k: v / n for k, v in sums.items()}, last_out, last_x, last_y


def validate(model, dataloader, device, args=None):
    """-> (avg G_loss, avg metrics, last Gx, last Fy or None, last x, last y), train.py:131-171."""
    model.eval()
    sums, n, total = {}, 0, 0.0
    last = (None, None, None, None)
    with torch.no_grad():
        for batch in dataloader:
            x, y = batch["x"].to(device, non_blocking=True), batch["y"].to(device, non_blocking=True)
            metrics = model.validation_step({"x": x, "y": y})
            gx, fy = metrics.pop("Gx"), metrics.pop("Fy", None)
            total += metrics["G_loss"]
            _accumulate(sums, metrics)
            n += 1
            last = (gx, fy, x, y)
    n = max(1, n)
    return (total / n, {k: v / n for k, v in sums.items()}, *last)


def _writer(log_dir, enabled):
    if not enabled:
        return None
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir=str(log_dir))
    except Exception as e:          # tensorboard missing: scalars are still printed
        print(f"TensorBoard disabled: {e}")
        return None


def main(args):
    from . import dist as vdist
    from . import plan
    from .utils import (load_checkpoint, load_pretrained_doubleae_to_cycleae, load_pretrained_doublevae_to_cyclevae,
                        save_checkpoint, truncate_tensorboard_events)
    arch = canonical_architecture(args.architecture)
    args.architecture = arch
    if arch in ("autoencoder", "vae") and args.source_modality != args.target_modality:
        raise ValueError("Source and target modalities should be the same for Autoencoder/VAE architectures.")
    src, tgt = DATASET_MODALITIES[args.dataset]
    args.source_modality = args.source_modality or src
    args.target_modality = args.target_modality or tgt
    if args.dataset == "summer2winter" and args.paired:
        print("WARNING: --paired flag is ignored for summer2winter dataset (inherently unpaired)")
        args.paired = False
    if args.pretrained_doubleae is not None and args.pretrained_doublevae is not None:
        raise ValueError("Cannot specify both --pretrained_doubleae and --pretrained_doublevae")
    if args.pretrained_doubleae is not None and arch not in ("cycleae", "cyclevae", "cycleaegan", "cyclevaegan"):
        raise ValueError(f"--pretrained_doubleae can only be used with Cycle architectures, not {arch}")
    if args.pretrained_doublevae is not None and arch not in ("cyclevae", "cyclevaegan"):
        raise ValueError(f"--pretrained_doublevae can only be used with CycleVAE or CycleVAEGAN architectures, not {arch}")
    if args.no_cuda or not torch.cuda.is_available():
        raise SystemExit("train.py: a CUDA device is required (B200 / sm_100a); the reference's CPU path is not "
                         "reimplemented (bench.py --impl reference times it)")
    rank, local, world = vdist.init_from_env("nccl")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    plan.set_precision(args.precision)
    main_rank = rank == 0

    # ---- run directory (train.py:396-421)
    if args.resume:
        ckpt_path = Path(args.resume)
        if not ckpt_path.exists():
            raise FileNotFoundError(f"No checkpoint found at {ckpt_path}")
        output_dir = ckpt_path.parent
    else:
        stamp = datetime.now().strftime("%m%d_%H%M")
        output_dir = Path(args.output_dir) / f"{arch}_{stamp}_{args.source_modality}_to_{args.target_modality}_{args.dataset}"
        if main_rank:
            output_dir.mkdir(parents=True, exist_ok=True)
            with open(output_dir / "args.json", "w") as f:
                json.dump(vars(args), f, indent=2)
    tb_dir = output_dir / "tensorboard"
    if args.resume and main_rank:
        truncate_tensorboard_events(tb_dir, torch.load(args.resume, map_location="cpu", weights_only=False)["epoch"])
    writer = _writer(tb_dir, main_rank)

    # ---- data (train.py:429-437)
    if args.dataset == "synthetic":
        train_loader = SyntheticPairs(args.batch_size, args.steps_per_epoch, args.image_size, device, rank, world,
                                      same_xy=(arch == "autoencoder"))
        test_loader = SyntheticPairs(args.batch_size, max(1, args.steps_per_epoch // 5), args.image_size, device, rank, world,
                                     same_xy=(arch == "autoencoder"), seed=8) if args.test_split > 0 else None
    else:
        from .Data_Manager import create_dataloaders
        train_loader, test_loader = create_dataloaders(args, device, rank, world)

    # ---- model, pretraining, optimizers, losses (train.py:440-469)
    torch.manual_seed(args.seed)
    model = create_model(arch, paired=args.paired, latent_dim=args.latent_dim).to(device)
    if args.pretrained_doubleae is not None:
        load_pretrained_doubleae_to_cycleae(model, args.pretrained_doubleae, device)
    if args.pretrained_doublevae is not None:
        load_pretrained_doublevae_to_cyclevae(model, args.pretrained_doublevae, device)
    model.configure_optimizers(lr=args.lr)
    model.configure_loss(lambda_kl=args.lambda_kl, lambda_gan=args.lambda_gan, lambda_identity=args.lambda_identity,
                         lambda_cycle=args.lambda_cycle, lambda_recon=args.lambda_recon)
    start_epoch = 0
    if args.resume:
        start_epoch = load_checkpoint(model, args.resume, device)[0] + 1
    if world > 1:
        vdist.broadcast_state(model)
        vdist.attach(model)
    torch.manual_seed(args.seed + 1 + rank)

    def report(tag, loss, comps):
        if main_rank:
            print(f"{tag} Loss: {loss:.4f}  " + "  ".join(f"{k}={v:.6f}" for k, v in sorted(comps.items())), flush=True)

    if test_loader is not None:
        loss0, comps0, *_ = validate(model, test_loader, device, args)
        report("Initial Test", loss0, comps0)
    best = float("inf")
    for epoch in range(start_epoch, args.epochs):
        if hasattr(train_loader, "set_epoch"):
            train_loader.set_epoch(epoch)
        t0 = time.time()
        train_loss, comps, _, _, _ = train_epoch(model, train_loader, device, args, writer=writer, epoch=epoch)
        dt = time.time() - t0
        if main_rank:
            print(f"\nEpoch {epoch + 1}/{args.epochs}  {args.batch_size * len(train_loader) / max(dt, 1e-9):.1f} img/s")
        report("Train", train_loss, comps)
        if writer is not None:
            writer.add_scalar("Loss/train", train_loss, epoch)
            for k, v in comps.items():
                writer.add_scalar(f"Loss_Components_train/{k}", v, epoch)
        if test_loader is not None and epoch % args.log_image_freq == 0:
            test_loss, tcomps, gx, fy, tx, ty = validate(model, test_loader, device, args)
            report("Test", test_loss, tcomps)
            if writer is not None:
                writer.add_scalar("Loss/test", test_loss, epoch)
                for k, v in tcomps.items():
                    writer.add_scalar(f"Loss_Components_test/{k}", v, epoch)
                writer.add_images(f"{args.source_modality}/test_x", tx[:4].clamp(0, 1), epoch)
                writer.add_images(f"{args.target_modality}/test_y", ty[:4].clamp(0, 1), epoch)
                writer.add_images(f"{args.target_modality}/test_Gx", gx[:4].clamp(0, 1), epoch)
                if fy is not None:
                    writer.add_images(f"{args.source_modality}/test_Fy", fy[:4].clamp(0, 1), epoch)
            if test_loss < best:
                best = test_loss
                if main_rank:
                    save_checkpoint(model, epoch, test_loss, args, output_dir / "best_model.pth")
        if (epoch + 1) % args.save_freq == 0 and main_rank:
            save_checkpoint(model, epoch, train_loss, args, output_dir / f"checkpoint_epoch_{epoch + 1}.pth")
    if writer is not None:
        writer.close()
    if main_rank:
        print(f"\nTraining completed. Models saved to {output_dir}")
    return model


if __name__ == "__main__":
    main(build_parser().parse_args())
