"""Run-to-run noise of a VAEGAN training step from identical state (atomics order only)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import Networks as N, plan
from oracle import ref_port as rp
plan.set_precision("bf16")
torch.manual_seed(1)
m = N.VAEGAN().cuda(); m.configure_optimizers(lr=2e-4); m.configure_loss()
batch = {k: v.cuda() for k, v in rp.synthetic_batch(1).items()}
m.training_step(batch)
sd = {k: v.clone() for k, v in m.state_dict().items()}
opt = m.save_optimizer_states()
import copy
outs = []
for trial in range(4):
    torch.manual_seed(2 + trial)
    m2 = N.VAEGAN().cuda(); m2.configure_optimizers(lr=2e-4); m2.configure_loss()
    m2.load_state_dict(sd); m2.load_optimizer_states(copy.deepcopy(opt))
    torch.manual_seed(9)
    outs.append(m2.training_step(batch))
torch.manual_seed(9)
outs.append(m.training_step(batch))
for k in outs[0]:
    print(f"{k:24s}", " ".join(f"{o[k]:10.5f}" for o in outs))
