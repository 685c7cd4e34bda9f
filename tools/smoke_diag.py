"""smoke() without the assert: prints every metric of the batch-1 CycleVAEGAN step next to the oracle's."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vcg_b200  # noqa
from oracle import ref_port as rp
from vcg_b200 import Networks as N
from vcg_b200 import lib, plan

torch.cuda.set_device(0)
lib.load()
plan.set_precision(os.environ.get("SMOKE_PREC", "bf16"))
torch.manual_seed(1234)
model = N.CycleVAEGAN(paired=False)
state = {k: v.clone() for k, v in model.state_dict().items()}
model = model.cuda()
model.configure_optimizers(lr=2e-4)
model.configure_loss(**rp.DEFAULT_LAMBDAS)
model.train()
batch = rp.synthetic_batch(int(os.environ.get("SMOKE_B", "1")))
N.set_eps_source(lambda shape, device: torch.randn(*shape).to(device))
torch.manual_seed(100)
ours = model.training_step({k: v.cuda() for k, v in batch.items()})
N.set_eps_source(None)
out = {k: float(v) for k, v in ours.items()}
if os.environ.get("SMOKE_ORACLE", "1") == "1":
    torch.manual_seed(100)
    ora = rp.RefModel("cyclevaegan", paired=False, state=state).training_step(batch)
    for k, v in ora.items():
        print(f"{k:22s} ours {out[k]:12.6f} oracle {v:12.6f} rel {abs(out[k]-v)/max(abs(v),1e-3):8.4f}")
else:
    print(json.dumps({k: round(v, 6) for k, v in out.items() if k.startswith(("D_loss", "d_", "loss_gan"))}))
