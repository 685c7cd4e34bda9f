"""The oracle (oracle/ref_port.py) against the golden vectors the REAL reference produced
(tests/golden/*.json, written by oracle/make_golden.py in the build container)."""
import json
import os

import pytest
import torch

from oracle import ref_port as rp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(tag):
    with open(os.path.join(GOLDEN, f"metrics_{tag}.json")) as f:
        return json.load(f)


def test_golden_files_cover_every_architecture():
    for arch in rp.ARCHS:
        g = gold(arch)
        assert g["arch"] == arch and len(g["steps"]) == 2 and "G_loss" in g["steps"][0]
    for arch in ("cycleae", "cyclevae", "cycleaegan", "cyclevaegan"):
        assert "loss_trans" in gold(arch + "_paired")["steps"][0] or "loss_identity" in gold(arch + "_paired")["steps"][0]


@pytest.mark.parametrize("arch", ["autoencoder", "cyclevaegan"])
def test_init_replay_matches_reference_checksums(arch):
    """constructor RNG replay: parameter checksums equal the reference's seeded construction."""
    g = gold(arch)
    torch.manual_seed(g["model_seed"])
    st = rp.init_state(arch, g["latent_dim"])
    n_params = sum(v.numel() for k, v in st.items() if not rp.is_buffer(k))
    assert n_params == g["n_params"]
    for k, (s, a) in g["init_checksum"].items():
        assert float(st[k].double().sum()) == pytest.approx(s, rel=1e-9, abs=1e-9)
        assert float(st[k].double().abs().sum()) == pytest.approx(a, rel=1e-9)


def test_oracle_reproduces_reference_training_metrics():
    """two Autoencoder training steps: the port must reproduce the real reference's metrics (same torch
    build => bit-exact; 1e-6 slack for other CPU thread counts)."""
    g = gold("autoencoder")
    torch.manual_seed(g["model_seed"])
    m = rp.RefModel("autoencoder", lr=g["lr"])
    batch = rp.synthetic_batch(g["batch"], seed=g["data_seed"], same_xy=True)
    for seed, ref in zip(g["eps_seeds"], g["steps"]):
        torch.manual_seed(seed)
        got = m.training_step(batch)
        assert set(got) == set(ref)
        for k in ref:
            assert got[k] == pytest.approx(ref[k], rel=2e-6), (k, got[k], ref[k])


def test_noise_floor_fixture_is_consistent():
    """fp64 oracle metrics recorded next to the reference's fp32 metrics: step-0 deviations are tiny."""
    nf = json.load(open(os.path.join(GOLDEN, "noise_floor.json")))
    for tag, rec in nf.items():
        g = gold(tag)
        for k, v in rec["steps_fp64"][0].items():
            ref = g["steps"][0][k]
            assert abs(v - ref) <= 2e-3 * max(abs(v), 1e-3), (tag, k, v, ref)


def test_losses_small_cases():
    """Losses.py:14-121 restated: hand-checkable values incl. the clamp edge of the KL term."""
    a = torch.tensor([[1.0, -2.0], [0.5, 0.5]])
    b = torch.tensor([[0.0, -2.0], [1.5, 0.0]])
    assert float(rp.l1(a, b)) == pytest.approx((1 + 0 + 1 + 0.5) / 4)
    d = torch.tensor([0.5, -1.0])
    tot, real, fake = rp.gan_loss_gen(d, d)
    assert float(real) == pytest.approx((0.25 + 1.0) / 2) and float(fake) == pytest.approx((0.25 + 4.0) / 2)
    assert float(tot) == pytest.approx(float(real) + float(fake))
    mu = torch.zeros(3)
    lv = torch.tensor([0.0, 20.0, -20.0])            # clamped to 10 / -10
    expect = -0.5 * ((1 + 0 - 0 - 1) + (1 + 10 - torch.exp(torch.tensor(10.0))) + (1 - 10 - torch.exp(torch.tensor(-10.0)))) / 3
    assert float(rp.kl_loss(mu, lv)) == pytest.approx(float(expect), rel=1e-6)
