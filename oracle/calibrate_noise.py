"""Measures the reference's OWN numerical noise so that end-to-end tolerances are anchored, not guessed.

For every architecture (batch 1, the golden seeds) it runs the oracle port in fp64 and records, per
metric and step, the fp64 value next to the fp32 value of the real reference (tests/golden/metrics_*).
For the GAN architectures it also runs the REAL reference (imported from /root/reference) under
torch.autocast(cpu, bfloat16) for one step: how far the reference's own bf16 execution moves each
metric.  Output: tests/golden/noise_floor.json.  Build-container only (needs /root/reference).

The parity tests then require, for metric m:
   fp32 mode: |ours - fp64| <= max(1e-5 |fp64|, 4 |ref_fp32 - fp64|)
   bf16 mode: |ours - fp64| <= max(2e-2 |fp64|, 2 |ref_bf16_autocast - fp64|)   (where measured)
"""
import argparse
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_port as rp  # noqa: E402
from oracle.make_golden import build_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--archs", default=",".join(rp.ARCHS))
    ap.add_argument("--bf16", default="aegan,vaegan,cycleaegan,cyclevaegan")
    ap.add_argument("--reuse", action="store_true", help="keep fp64 values already in noise_floor.json")
    args = ap.parse_args()
    gdir = os.path.join(ROOT, "tests", "golden")
    out_path = os.path.join(gdir, "noise_floor.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    torch.set_num_threads(os.cpu_count())
    for arch in args.archs.split(","):
        for paired in ([False, True] if arch.startswith("cycle") else [False]):
            tag = arch + ("_paired" if paired else "")
            gold = json.load(open(os.path.join(gdir, f"metrics_{tag}.json")))
            if args.reuse and "steps_fp64" in out.get(tag, {}) and arch not in args.bf16.split(","):
                continue
            t0 = time.time()
            torch.manual_seed(gold["model_seed"])
            st = rp.init_state(arch, gold["latent_dim"])
            ora = rp.RefModel(arch, gold["latent_dim"], paired, state=st, dtype=torch.float64, lr=gold["lr"],
                              eps_source=lambda std: torch.randn(std.shape).to(std.dtype))
            batch = rp.synthetic_batch(gold["batch"], seed=gold["data_seed"], same_xy=(arch == "autoencoder"),
                                       dtype=torch.float64)
            rec = out.setdefault(tag, {})
            steps64 = rec["steps_fp64"] if (args.reuse and "steps_fp64" in rec) else []
            for s, seed in enumerate(gold["eps_seeds"]):
                if len(steps64) == len(gold["eps_seeds"]):
                    break
                torch.manual_seed(seed)
                steps64.append(ora.training_step(batch))
            rec["steps_fp64"] = steps64
            rec["ref_fp32_rel_dev"] = [
                {k: abs(gold["steps"][s][k] - steps64[s][k]) / max(abs(steps64[s][k]), 1e-12) for k in steps64[s]}
                for s in range(len(steps64))]
            print(f"[noise] {tag}: fp64 done in {time.time() - t0:.1f}s; worst fp32 dev step0 "
                  f"{max(rec['ref_fp32_rel_dev'][0].values()):.2e} step1 {max(rec['ref_fp32_rel_dev'][1].values()):.2e}",
                  flush=True)
            if arch in args.bf16.split(","):
                t0 = time.time()
                torch.manual_seed(gold["model_seed"])
                ref = build_reference(arch, gold["latent_dim"], paired)
                ref.configure_optimizers(lr=gold["lr"])
                ref.configure_loss(**gold["lambdas"])
                ref.train()
                b32 = rp.synthetic_batch(gold["batch"], seed=gold["data_seed"], same_xy=(arch == "autoencoder"))
                devs = []
                for s, seed in enumerate(gold["eps_seeds"]):
                    torch.manual_seed(seed)
                    with torch.autocast("cpu", dtype=torch.bfloat16):
                        m = ref.training_step(b32)
                    devs.append({k: abs(m[k] - steps64[s][k]) / max(abs(steps64[s][k]), 1e-12) for k in m})
                rec["ref_bf16_rel_dev"] = devs[0]
                rec["ref_bf16_rel_dev_steps"] = devs
                print(f"[noise] {tag}: reference under bf16 autocast in {time.time() - t0:.1f}s; worst dev step0 "
                      f"{max(devs[0].values()):.2f} step1 {max(devs[1].values()):.2f}", flush=True)
            with open(out_path, "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
