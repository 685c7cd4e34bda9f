"""Checkpoint interop with the REAL reference (build container only: needs /root/reference).

TEST INFRASTRUCTURE (see oracle/ref_port.py).  Both directions of the on-disk contract of utils.py:17-54:
  1. the reference's own save_checkpoint() of a reference model that took one training step  ->  our load_checkpoint()
     into our model + FusedAdam (CPU tensors; no kernel runs): state_dict and Adam state bit-equal;
  2. our save_checkpoint()  ->  the reference's load_checkpoint() into a fresh reference model: bit-equal again;
  3. the reference's Double-VAE -> Cycle-VAE transfer on a file we wrote equals ours on a file it wrote.
usage: python oracle/check_interop.py [--arch vaegan]"""
import argparse
import importlib.util
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import ref_port as rp  # noqa: E402


def ref_utils():
    spec = importlib.util.spec_from_file_location("_reference_utils", "/root/reference/utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vaegan")
    a = ap.parse_args()
    net, origin = ref_loader.load()
    assert net is not None, origin
    ru = ref_utils()
    import vcg_b200  # noqa: F401
    from vcg_b200 import Networks as N
    from vcg_b200 import utils as ou
    cls = {"vaegan": "VAEGAN", "autoencoder": "Autoencoder", "cyclevae": "CycleVAE"}[a.arch]
    torch.manual_seed(11)
    ref = getattr(net, cls)()
    ref.configure_optimizers(lr=2e-4)
    ref.configure_loss(**rp.DEFAULT_LAMBDAS)
    ref.train()
    ref.training_step(rp.synthetic_batch(1))
    tmp = tempfile.mkdtemp(prefix="vcg_interop_")
    f1, f2 = os.path.join(tmp, "by_reference.pth"), os.path.join(tmp, "by_ours.pth")
    ns = argparse.Namespace(architecture=a.arch, paired=True, lr=2e-4)
    ru.save_checkpoint(ref, 7, 0.5, ns, f1)
    ours = getattr(N, cls)()
    ours.configure_optimizers(lr=2e-4)
    assert ou.load_checkpoint(ours, f1, "cpu") == (7, 0.5)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), (k1, k2)
    so, sr = ours.save_optimizer_states(), ref.save_optimizer_states()
    assert set(so) == set(sr)
    for name in sr:
        for i, st in sr[name]["state"].items():
            for key in ("step", "exp_avg", "exp_avg_sq"):
                assert torch.equal(torch.as_tensor(st[key]), torch.as_tensor(so[name]["state"][i][key])), (name, i, key)
    print(f"[interop] reference file -> ours: {len(ours.state_dict())} state entries and the Adam state of {list(sr)} bit-equal")
    ou.save_checkpoint(ours, 8, 0.25, ns, f2)
    ref2 = getattr(net, cls)()
    ref2.configure_optimizers(lr=2e-4)
    assert ru.load_checkpoint(ref2, f2, "cpu") == (8, 0.25)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), ref2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    s2 = ref2.save_optimizer_states()
    for name in sr:
        for i, st in sr[name]["state"].items():
            assert torch.equal(st["exp_avg_sq"], s2[name]["state"][i]["exp_avg_sq"])
    print("[interop] our file -> reference load_checkpoint(): bit-equal")
    # pretrain transfer, both implementations on both files
    torch.manual_seed(12)
    dv = net.DoubleVariationalAutoencoder()
    dv.configure_optimizers(lr=2e-4)
    f3 = os.path.join(tmp, "doublevae.pth")
    ru.save_checkpoint(dv, 0, 0.0, ns, f3)
    c_ref, c_ours = net.CycleVAE(), N.CycleVAE()
    ru.load_pretrained_doublevae_to_cyclevae(c_ref, f3, "cpu")
    ou.load_pretrained_doublevae_to_cyclevae(c_ours, f3, "cpu")
    for (k1, v1), (k2, v2) in zip(c_ref.state_dict().items(), c_ours.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    print("[interop] Double-VAE -> Cycle-VAE transfer: ours == reference's on the same file")
    for f in (f1, f2, f3):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
