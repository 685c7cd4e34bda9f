"""Uninitialised-read detector: poison the caching allocator's free memory with NaN before a training step."""
import os, sys, copy
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import Networks as N, plan
from oracle import ref_port as rp
plan.set_precision("bf16")
arch = sys.argv[1] if len(sys.argv) > 1 else "VAEGAN"
torch.manual_seed(1)
m = getattr(N, arch)().cuda(); m.configure_optimizers(lr=2e-4); m.configure_loss()
batch = {k: v.cuda() for k, v in rp.synthetic_batch(2).items()}
outs = []
for trial in range(3):
    if trial > 0:
        junk = torch.full((2_000_000_000,), float("nan"), dtype=torch.bfloat16, device="cuda")   # 4 GB of NaN
        del junk
    torch.manual_seed(9)
    outs.append(m.training_step(batch))
    bad = [k for k, v in m.state_dict().items() if not torch.isfinite(v.float()).all()]
    print("trial", trial, "non-finite params:", bad[:5])
for k in outs[0]:
    print(f"{k:24s}", " ".join(f"{o[k]:12.5f}" for o in outs))
