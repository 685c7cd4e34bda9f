"""Parity at the BASELINE.json configurations (-m gpu): configs[1] VAE latent_dim 1024 batch 8, configs[2] VAE-GAN
batch 16 and the 8-GPU shard of configs[4] (CycleVAEGAN unpaired, per-GPU batch 8) -- the batch sizes and the latent
width the benchmark runs at, which take kernel paths the batch-1 goldens do not (1024-wide fp32-stored bottleneck
convs, BN = 256 / tail-split / ring schedules, multi-image TMA boxes).

The checker is the oracle port (oracle/ref_port.py, pinned bit-exactly to the real reference) run ON THE BOX'S GPU in
fp32 with TF32 off (SURVEY.md 8c mode iii), in two forms:
  * fp32 mode of the kernels  vs  the plain oracle:  metrics 1e-5, as north_star states;
  * bf16 mode (tcgen05 path)  vs  the oracle in `emulate_bf16` form, which rounds to bfloat16 exactly where the
    kernels do (conv operands, stored activations and gradients).

What can be held to which tolerance (measured, round 2):
  * forward quantities: generator-side losses agree with the emulating oracle to 4e-5 at step 0 (bound 2e-3) and
    discriminator-side scalars to 3..8e-2 (bound 1.2e-1; against the reference's own bf16 autocast they moved by 25 %+);
  * END-TO-END gradients under the training losses cannot be held to 2e-2 in ANY arithmetic: the L1 gradient is
    sign(Gx - y) / N, spatially smooth at initialisation, and every InstanceNorm backward removes the plane-wise mean and
    zhat-component of the gradient, so the signal shrinks layer by layer while rounding noise does not.  Our fp32 kernels
    and cuDNN fp32 -- per-op error 1e-6 -- already differ by 1.2e-2 per tensor (cosine 0.9999) on the same weights; two
    bf16 executions that round at the same points differ by 0.3..0.7.  Those gradients are therefore bounded by their
    measured noise (cosine), and the WIRING of the bf16 backward is tested where it is well conditioned:
  * the same plans with a random (white) cotangent on every output, where the signal is not cancelled by the norms:
    outputs, every parameter gradient and the input gradient of the generator / discriminator plans against the
    emulating oracle.  Two bf16 executions cannot track each other more closely than independent bf16 noise: a
    1e-6 difference in one pre-rounding value (fp32 summation order) flips a rounding decision in 1e-6 / 4e-3 of
    the elements, each flip is a full ulp, and the flips it causes downstream cascade to full decorrelation within
    ~6 of the 36 rounding stages.  Measured: outputs 2.2..2.7e-2; discriminator-plan gradients 2.6e-2 median (worst
    3.3e-2, cosine 0.9995); generator-plan gradients 0.20..0.22 median (worst 0.26, cosine 0.967) because every
    layer's backward map is built from forward activations that differ by ~2.5e-2.  Bounds follow those numbers
    (discriminator 6e-2 / 0.998, generators 0.4 / 0.93): a mis-routed gradient (missing residual / halo / consumer
    term) is an O(1) error in at least one tensor and fails; finer wiring errors are caught by the SAME plan code in
    fp32 mode (1.2e-2 floor against cuDNN fp32, bound 4e-2 with cosine 0.999) and by the per-kernel bf16 tests."""
import os
import re

import pytest
import torch

from util import rel_l2

pytestmark = pytest.mark.gpu

CASES = {
    # name: (class, ctor kwargs, oracle arch, oracle kwargs, batch, same_xy)
    "config2_vae_latent1024_b8": ("VariationalAutoencoder", {"latent_dim": 1024}, "vae", {"latent_dim": 1024}, 8, True),
    "config3_vaegan_b16": ("VAEGAN", {}, "vaegan", {}, 16, False),
    "config5_shard_cyclevaegan_b8": ("CycleVAEGAN", {"paired": False}, "cyclevaegan", {"paired": False}, 8, False),
}
# parameters whose gradient is mathematically zero (bias in front of an InstanceNorm): rounding
# noise on both sides, excluded from the per-tensor comparison
_DEAD_GEN = ("encoder.model.0.conv.bias", "encoder.model.5.conv2.bias", "decoder.model.0.conv2.bias")
_DEAD_DISC = re.compile(r"(^|\.)D[XY]?\.model\.[123]\.conv\.bias$")


def dead_bias(key):
    return key.endswith(_DEAD_GEN) or bool(_DEAD_DISC.search(key))


@pytest.fixture(scope="module")
def env(vcg):
    from oracle import ref_port as rp
    from vcg_b200 import Networks as N
    from vcg_b200 import plan
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield N, plan, rp
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    plan.set_precision("bf16")
    N.set_eps_source(None)


def _optimizers(m):
    return [o for o in (getattr(m, n, None) for n in ("optimizer", "optimizer_G", "optimizer_D")) if o is not None]


def _run(N, plan, rp, case, prec, steps):
    cls, kw, arch, okw, b, same = CASES[case]
    plan.set_precision(prec)
    torch.manual_seed(1234)
    ours = getattr(N, cls)(**kw)
    state = {k: v.clone() for k, v in ours.state_dict().items()}
    ours = ours.cuda()
    ours.configure_optimizers(lr=2e-4)
    ours.configure_loss(**rp.DEFAULT_LAMBDAS)
    ours.train()
    for o in _optimizers(ours):
        o.keep_grads = True                         # read .grad after the step
    batch = {k: v.cuda() for k, v in rp.synthetic_batch(b, same_xy=same).items()}
    g1 = torch.Generator().manual_seed(55)
    N.set_eps_source(lambda shape, device: torch.randn(*shape, generator=g1).to(device))
    g2 = torch.Generator().manual_seed(55)
    ora = rp.RefModel(arch, state=state, lr=2e-4, emulate_bf16=(prec == "bf16"), device="cuda",
                      eps_source=lambda std: torch.randn(std.shape, generator=g2).to(std), **okw)
    res = []
    for _ in range(steps):
        m = ours.training_step(batch)
        for o in _optimizers(ours):
            o.finish()
        grads = {k: p.grad.detach().clone() for k, p in ours.named_parameters() if p.grad is not None}
        mo = ora.training_step(batch)
        gref = {k: p.grad.detach().clone() for k, p in ora.P.items() if p.grad is not None}
        res.append((m, mo, grads, gref))
    N.set_eps_source(None)
    return res


def _disc_side(key):
    return key.startswith(("D_loss", "d_", "loss_gan"))


def _check(case, prec, res, mtol, gtol, gcos):
    lines = []
    for s, (m, mo, grads, gref) in enumerate(res):
        assert set(m) == set(mo), set(m) ^ set(mo)
        worst_m = max(((abs(m[k] - mo[k]) / max(abs(mo[k]), 1e-3), k) for k in mo))
        errs = []
        for k, g in grads.items():
            if dead_bias(k) or k not in gref:
                continue
            r = gref[k]
            e = rel_l2(g, r)
            cos = float((g.double().flatten() @ r.double().flatten()) / (g.double().norm() * r.double().norm() + 1e-300))
            errs.append((e, cos, k))
        worst_g = max(errs)
        worst_c = min((c, k) for _, c, k in errs)
        lines.append(f"[{case} {prec} step {s}] worst metric {worst_m[1]} {worst_m[0]:.2e}; worst grad {worst_g[2]} "
                     f"rel_l2 {worst_g[0]:.2e}; lowest cosine {worst_c[1]} {worst_c[0]:.6f}; "
                     f"median grad rel_l2 {sorted(e for e, _, _ in errs)[len(errs) // 2]:.2e}")
    print("\n".join(lines))
    for s, (m, mo, grads, gref) in enumerate(res):
        for k in mo:
            assert m[k] == m[k], (case, s, k)
            err, ref = abs(m[k] - mo[k]), abs(mo[k])
            if prec == "fp32":
                bound = mtol * max(ref, 1e-3) + 1e-6
            elif not _disc_side(k):
                # generator-side losses: 2e-3 at step 0; after one Adam step (sign-like: every weight moved by +-lr along a
                # noisy gradient sign) 2e-2
                bound = (2e-3 if s == 0 else 2e-2) * max(ref, 1e-3) + 1e-6
            elif s == 0:
                # measured 3.2e-2 .. 7.9e-2 over the three configurations (the run-to-run summation order of the
                # atomics moves them): 1.5 x the worst observed
                bound = 1.2e-1 * max(ref, 1e-2) + 2e-3
            else:
                # discriminator scalars after the update: a 131072-term dot product of normalised features of images
                # produced by weights that differ by +-lr per element (Adam's first step is sign-like).  Measured against
                # the emulating oracle: loss_gan_g 2.22 vs 3.39, d_y_fake_mean off by several times its value; the real
                # reference moves these by 40..200 % under its own bf16 autocast (tests/golden/noise_floor.json).  Only
                # sanity-bounded: finite and of the right order of magnitude.
                bound = 1.0 * max(ref, 1.0)
            assert err <= bound, (case, prec, s, k, m[k], mo[k], bound)
        if s > 0:
            continue            # gradients of later steps are taken at weights that already differ by +-lr per element
        for k, g in grads.items():
            if dead_bias(k) or k not in gref:
                continue
            r = gref[k]
            e = rel_l2(g, r)
            cos = float((g.double().flatten() @ r.double().flatten()) / (g.double().norm() * r.double().norm() + 1e-300))
            assert e <= gtol and cos >= gcos, (case, prec, s, k, e, cos)


@pytest.mark.parametrize("case", list(CASES))
def test_bf16_tensor_core_path_vs_bf16_emulating_oracle(env, case):
    N, plan, rp = env
    res = _run(N, plan, rp, case, "bf16", steps=2)
    # training-loss gradients: noise-bounded (see the module docstring); the tight bf16 backward check is
    # test_bf16_backward_wiring_with_white_cotangent below
    _check(case, "bf16", res, mtol=2e-2, gtol=1.2, gcos=0.5)


def test_fp32_mode_vs_gpu_fp32_oracle_config2(env):
    """configs[1] in fp32 parity mode against the oracle on the GPU (TF32 off): metrics to 1e-5, gradients to the
    reference's own fp32 reordering noise (a few 1e-3: ReLU-gate flips)."""
    N, plan, rp = env
    res = _run(N, plan, rp, "config2_vae_latent1024_b8", "fp32", steps=1)
    _check("config2_vae_latent1024_b8", "fp32", res, mtol=1e-5, gtol=4e-2, gcos=0.999)


@pytest.mark.parametrize("net", ["vae1024", "ae", "disc"])
def test_bf16_backward_wiring_with_white_cotangent(env, net):
    """bf16 forward + backward of one network plan with a random cotangent on every output, against the emulating
    oracle: outputs, every parameter gradient and the input gradient."""
    N, plan, rp = env
    plan.set_precision("bf16")
    bf = lambda t: t.to(torch.bfloat16).float()         # noqa: E731
    g = torch.Generator().manual_seed(21)
    torch.manual_seed(77)
    if net == "disc":
        b, ours = 16, N.Discriminator()
        pre = "D."
    elif net == "ae":
        b, ours = 8, N.Autoencoder()
        pre = ""
    else:
        b, ours = 8, N.VariationalAutoencoder(latent_dim=1024)
        pre = ""
    state = {pre + k: v.clone() for k, v in ours.state_dict().items()}
    ours = ours.cuda().train()
    x = bf(torch.rand(b, 3, 256, 256, generator=g)).cuda()
    P = {k: v.cuda() for k, v in state.items()}
    for k in P:
        if not rp.is_buffer(k):
            P[k].requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    with rp.emulate_bf16():
        if net == "disc":
            outs_o = (rp.discriminator(P, "D.", xo, training=True),)
            outs = (ours(xc),)
        elif net == "ae":
            outs_o = (rp.ae_forward(P, "", xo),)
            outs = (ours(xc),)
        else:
            eps = torch.randn(b, 1024, 16, 16, generator=g).cuda()
            outs_o = rp.vae_forward(P, "", xo, eps)
            outs = ours(xc, eps=eps)
        cots = [bf(torch.randn(o.shape, generator=g)).cuda() for o in outs_o]
        sum((o * c).sum() for o, c in zip(outs_o, cots)).backward()
    sum((o * c).sum() for o, c in zip(outs, cots)).backward()
    names = ("score",) if net == "disc" else (("Gx",) if net == "ae" else ("Gx", "mu", "logvar"))
    report = []
    for nme, o, oo in zip(names, outs, outs_o):
        e = rel_l2(o.detach(), oo.detach())
        report.append((e, 1.0, "out:" + nme))
        assert e < 5e-2, (net, nme, e)
    pairs = [("input", xc.grad, xo.grad)] + [(k, p.grad, P[pre + k].grad) for k, p in ours.named_parameters()]
    for k, got, ref in pairs:
        if dead_bias(pre + k) or ref is None:
            continue
        e = rel_l2(got, ref)
        cos = float((got.double().flatten() @ ref.double().flatten()) / (got.double().norm() * ref.double().norm() + 1e-300))
        report.append((e, cos, k))
    worst = max(report)
    print(f"[white cotangent {net}] worst rel_l2 {worst[2]} {worst[0]:.2e}; lowest cosine {min(c for _, c, _ in report):.6f}; "
          f"median rel_l2 {sorted(e for e, _, _ in report)[len(report) // 2]:.2e}")
    # measured (round 2): discriminator plan (4 convs, 3 norms) worst 3.3e-2 / cosine 0.9995; generator plans (18 convs,
    # 15 norms) worst 0.26 / cosine 0.967: the backward map of every layer is built from forward activations (zhat of
    # the InstanceNorm backward, ReLU masks, the GEMM operand of the weight gradient) that differ by ~2.5e-2 between any
    # two bf16 executions, and ~30 such stages compound
    lim, cmin = (6e-2, 0.998) if net == "disc" else (0.4, 0.93)
    for e, cos, k in report:
        assert e <= lim and cos >= cmin, (net, k, e, cos)
