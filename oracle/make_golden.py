"""Pins the oracle (oracle/ref_port.py) against the REAL reference and writes tests/golden/.

Runs only in the build container (needs /root/reference); the fixtures it writes are
committed so the GPU box never needs the reference.  Checks, per architecture:
  1. constructor RNG replay: init_state(arch) == reference state_dict(), bit for bit;
  2. two training steps on a seeded synthetic batch: every metric equal to the
     reference's own training_step (CPU fp32, same torch build => bit-exact);
  3. state_dict after the two steps equal.
Writes tests/golden/metrics_<arch>.json (+ parameter checksums) and a small npz with
sub-sampled forward tensors for the CycleVAEGAN case.

usage: python oracle/make_golden.py [--archs a,b,...] [--batch 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_port as rp  # noqa: E402

REF_CLASS = {"autoencoder": "Autoencoder", "vae": "VariationalAutoencoder", "aegan": "AEGAN",
             "vaegan": "VAEGAN", "cycleae": "CycleAE", "cyclevae": "CycleVAE",
             "cycleaegan": "CycleAEGAN", "cyclevaegan": "CycleVAEGAN",
             "doubleae": "DoubleAutoencoder", "doublevae": "DoubleVariationalAutoencoder"}


def build_reference(arch, latent, paired):
    import Networks  # the reference module
    cls = getattr(Networks, REF_CLASS[arch])
    kw = {}
    if arch in ("vae", "vaegan", "cyclevae", "cyclevaegan", "doublevae"):
        kw["latent_dim"] = latent
    if arch.startswith("cycle"):
        kw["paired"] = paired
    return cls(**kw)


def checksum(state):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in state.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--archs", default=",".join(rp.ARCHS + rp.DOUBLE_ARCHS))
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--latent", type=int, default=64)
    args = ap.parse_args()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for arch in args.archs.split(","):
        for paired in ([False, True] if arch.startswith("cycle") else [False]):
            t0 = time.time()
            torch.manual_seed(1234)
            ref = build_reference(arch, args.latent, paired)
            ref_sd = ref.state_dict()
            torch.manual_seed(1234)
            st = rp.init_state(arch, args.latent)
            assert list(st.keys()) == list(ref_sd.keys()), (arch, "state_dict keys/order differ")
            for k in st:
                assert torch.equal(st[k], ref_sd[k]), (arch, k, "init differs")
            init_sum = checksum(st)
            ora = rp.RefModel(arch, args.latent, paired, state=st, lr=2e-4)
            ref.configure_optimizers(lr=2e-4)
            ref.configure_loss(**rp.DEFAULT_LAMBDAS)
            ref.train()
            batch = rp.synthetic_batch(args.batch, same_xy=(arch == "autoencoder"))
            steps = []
            for s in range(args.steps):
                torch.manual_seed(100 + s)
                m_ref = ref.training_step(batch)
                torch.manual_seed(100 + s)
                m_ora = ora.training_step(batch)
                assert set(m_ref) == set(m_ora), (arch, set(m_ref) ^ set(m_ora))
                for k in m_ref:
                    assert m_ref[k] == m_ora[k], (arch, s, k, m_ref[k], m_ora[k])
                steps.append(m_ref)
            sd_ref, sd_ora = ref.state_dict(), ora.state_dict()
            for k in sd_ref:
                assert torch.equal(sd_ref[k], sd_ora[k]), (arch, k, "post-step state differs")
            tag = arch + ("_paired" if paired else "")
            rec = {"arch": arch, "paired": paired, "latent_dim": args.latent, "batch": args.batch,
                   "model_seed": 1234, "data_seed": 7, "eps_seeds": [100 + s for s in range(args.steps)],
                   "lr": 2e-4, "lambdas": rp.DEFAULT_LAMBDAS, "torch": torch.__version__,
                   "threads": torch.get_num_threads(), "steps": steps,
                   "n_params": int(sum(v.numel() for k, v in sd_ref.items() if not rp.is_buffer(k))),
                   "init_checksum": {k: init_sum[k] for k in list(init_sum)[:4] + list(init_sum)[-4:]},
                   "final_checksum": {k: v for k, v in list(checksum(sd_ref).items())[:4]}}
            with open(os.path.join(out_dir, f"metrics_{tag}.json"), "w") as f:
                json.dump(rec, f, indent=1)
            print(f"[golden] {tag}: port == reference over {args.steps} steps "
                  f"({time.time() - t0:.1f}s)  G_loss={steps[0]['G_loss']:.6f}", flush=True)


if __name__ == "__main__":
    main()
