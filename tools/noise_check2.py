import os, sys, copy
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
def val(model):
    torch.manual_seed(9)
    o = model.validation_step(batch)
    return {k: v for k, v in o.items() if not torch.is_tensor(v)}
outs = [val(m), val(m)]
models = []
for trial in range(3):
    torch.manual_seed(2 + trial)
    m2 = N.VAEGAN().cuda(); m2.configure_optimizers(lr=2e-4); m2.configure_loss()
    m2.load_state_dict(sd)
    models.append(m2)
    outs.append(val(m2)); outs.append(val(m2))
for k in outs[0]:
    print(f"{k:24s}", " ".join(f"{o[k]:10.5f}" for o in outs))
# weight equality
for k, v in models[0].state_dict().items():
    if not torch.equal(v, sd[k]): print("state differs", k)
