"""Throughput of the other BASELINE.json configurations on one B200 (CUDA-graph replay, device-resident batches).
usage: python tools/bench_configs.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import Networks as N, plan
from vcg_b200.graph import GraphedStep

plan.set_precision("bf16")
CONFIGS = [  # name, constructor, batch, same_xy
    ("config 1 shape: Autoencoder depth->depth B=4", lambda: N.Autoencoder(), 4, True),
    ("config 2: VAE latent_dim 1024 lambda_kl 1e-5 B=8", lambda: N.VariationalAutoencoder(latent_dim=1024), 8, True),
    ("config 3: VAE-GAN B=16", lambda: N.VAEGAN(), 16, False),
    ("config 4 (per-GPU share at 2 GPUs): Cycle-VAE unpaired B=16", lambda: N.CycleVAE(paired=False), 16, False),
    ("config 4 (whole batch on one GPU): Cycle-VAE unpaired B=32", lambda: N.CycleVAE(paired=False), 32, False),
]
for name, ctor, b, same in CONFIGS:
    torch.manual_seed(1234)
    m = ctor().cuda()
    m.configure_optimizers(lr=2e-4)
    m.configure_loss(lambda_kl=1e-5, lambda_gan=1.0, lambda_identity=5.0, lambda_cycle=10.0, lambda_recon=1.0)
    m.train()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(b, 3, 256, 256, generator=g).cuda()
    y = x if same else torch.rand(b, 3, 256, 256, generator=g).cuda()
    step = GraphedStep(m, {"x": x, "y": y}, warmup=2)
    for _ in range(3):
        out = step({"x": x, "y": y})
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    steps = 10
    for _ in range(steps):
        out = step({"x": x, "y": y})
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    print(json.dumps({"config": name, "batch": b, "ms_per_step": round(ms, 3), "img_per_s": round(b / ms * 1e3, 1),
                      "G_loss": round(float(out.get("G_loss", out.get("total_loss", 0.0))), 4)}), flush=True)
    del m, step
    torch.cuda.empty_cache()
