"""test.py -- the reference's inference entry (its test.py:31-142, 284-314) on the B200 forward kernels.

Finds finished runs under a runs directory (args.json + best_model.pth, the layout train.py writes), loads a checkpoint
in the reference's format into the matching model in eval() mode -- discriminators then use the stale spectral-norm
sigma exactly like the reference -- and runs the forward pass only (plans are executed without saving activations
under torch.no_grad()).  The reference's matplotlib comparison figures (test.py:345-605) are replaced by plain PNG
strips x | Gx | y written with PIL; figure styling is outside the accelerated path."""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path

import torch

from .train import canonical_architecture, create_model as _create_model


def discover_runs(runs_dir="runs"):
    """-> [{'name', 'path', 'architecture', 'args', 'checkpoint'}] for every run directory holding args.json and a
    best_model.pth (test.py:31-70 of the reference)."""
    found = []
    root = Path(runs_dir)
    if not root.is_dir():
        return found
    for d in sorted(p for p in root.iterdir() if p.is_dir()):
        a, ck = d / "args.json", d / "best_model.pth"
        if a.exists() and ck.exists():
            with open(a) as f:
                args = json.load(f)
            found.append({"name": d.name, "path": d, "architecture": args.get("architecture", "unknown"), "args": args,
                          "checkpoint": ck})
    return found


def create_model(architecture, paired=True, latent_dim=64):
    return _create_model(architecture, paired=paired, latent_dim=latent_dim)


def load_model_for_inference(architecture, checkpoint_path, device):
    """Model in eval mode from a reference-format checkpoint; `paired` (and latent_dim, when recorded) come from the
    args stored in the file (test.py:110-142)."""
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    saved = ckpt.get("args", {}) or {}
    model = create_model(architecture, paired=saved.get("paired", True), latent_dim=saved.get("latent_dim", 64))
    model.load_state_dict(ckpt["model_state_dict"])
    model = model.to(device).eval()
    epoch, loss = ckpt.get("epoch", "unknown"), ckpt.get("loss", "unknown")
    print(f"  Loaded {architecture} from epoch {epoch}" + (f" (loss: {loss:.4f})" if isinstance(loss, float) else ""))
    return model


def run_inference(model, batch, architecture, device, unpaired=False):
    """-> (output, x, y): the first output of forward() -- Gx by the reference's convention (Networks.py:16) -- for any
    architecture (test.py:284-314; legacy unpaired batches use the keys 'A' / 'B')."""
    from . import Networks as N
    model.eval()
    with torch.no_grad():
        kx, ky = ("A", "B") if (unpaired or "A" in batch) else ("x", "y")
        x, y = batch[kx].to(device), batch[ky].to(device)
        out = model(x) if isinstance(model, (N.Autoencoder, N.VariationalAutoencoder)) else model(x, y)
        return (out[0] if isinstance(out, (tuple, list)) else out), x, y


def save_strip(path, x, out, y, max_samples=8):
    """x | output | y rows as one PNG (values clamped to [0, 1], the ToTensor range of the datasets)"""
    from PIL import Image
    rows = []
    for i in range(min(max_samples, x.shape[0])):
        rows.append(torch.cat([t[i].detach().float().clamp(0, 1).cpu() for t in (x, out, y)], dim=2))
    img = (torch.cat(rows, dim=1).permute(1, 2, 0) * 255).round().to(torch.uint8).numpy()
    Image.fromarray(img).save(path)


def evaluate_models(args):
    device = torch.device("cuda")
    runs = discover_runs(args.runs_dir)
    if not runs:
        print(f"No runs with args.json + best_model.pth under {args.runs_dir}")
        return []
    os.makedirs(args.output_dir, exist_ok=True)
    done = []
    for run in runs:
        arch = canonical_architecture(run["architecture"])
        model = load_model_for_inference(arch, run["checkpoint"], device)
        ra = argparse.Namespace(**{**vars(args), **{k: v for k, v in run["args"].items() if k in
                                                     ("dataset", "data_dir", "source_modality", "target_modality", "image_size",
                                                      "paired", "test_split", "seed")}})
        ra.batch_size, ra.num_workers, ra.cuda_graph = args.num_samples, 0, False
        if ra.dataset == "synthetic":
            from .train import SyntheticPairs
            loader = SyntheticPairs(args.num_samples, 1, ra.image_size, device, same_xy=(arch == "autoencoder"), seed=8)
        else:
            from .Data_Manager import create_dataloaders
            loader = create_dataloaders(ra, device)[1]
        batch = next(iter(loader))
        out, x, y = run_inference(model, batch, arch, device)
        path = os.path.join(args.output_dir, f"{run['name']}.png")
        save_strip(path, x, out, y, args.num_samples)
        print(f"  wrote {path}")
        done.append(path)
    return done


def build_parser():
    p = argparse.ArgumentParser(description="Evaluate trained VAE-CycleGAN models (forward pass on B200 kernels)")
    p.add_argument("--runs_dir", type=str, default="runs")
    p.add_argument("--output_dir", type=str, default="test_results")
    p.add_argument("--num_samples", type=int, default=8)
    return p


if __name__ == "__main__":
    evaluate_models(build_parser().parse_args())
