"""Host-side logic without a GPU: module surface, state_dict contract, seeded construction, plan graphs,
CLI aliases, optimizer interface, and the loud failure when no CUDA device is present."""
import pytest
import torch

from oracle import ref_port as rp


@pytest.fixture(scope="module")
def N(vcg):
    from vcg_b200 import Networks
    return Networks


CLS = {"autoencoder": "Autoencoder", "vae": "VariationalAutoencoder", "aegan": "AEGAN", "vaegan": "VAEGAN",
       "cycleae": "CycleAE", "cyclevae": "CycleVAE", "cycleaegan": "CycleAEGAN", "cyclevaegan": "CycleVAEGAN"}


@pytest.mark.parametrize("arch", ["autoencoder", "vaegan", "cyclevaegan"])
def test_state_dict_contract_and_seeded_init(N, arch):
    """keys, order, shapes AND values of a seeded construction equal the reference's (via the oracle's
    constructor replay, which oracle/make_golden.py pins bit-exactly to the real reference)."""
    torch.manual_seed(1234)
    sd = getattr(N, CLS[arch])().state_dict()
    torch.manual_seed(1234)
    ref = rp.init_state(arch)
    assert list(sd) == list(ref)
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == torch.float32
        assert torch.equal(sd[k], ref[k]), k
    if arch == "cyclevaegan":
        assert len(sd) == 96
        assert tuple(sd["G.encoder.model.1.conv.weight"].shape) == (128, 256, 3, 3)
        assert tuple(sd["DY.model.4.weight_orig"].shape) == (1, 512, 16, 16)
        assert tuple(sd["DY.model.4.weight_u"].shape) == (1,) and tuple(sd["DY.model.4.weight_v"].shape) == (131072,)


def test_reference_class_names_and_signatures(N):
    for name in ("CaSb", "D", "R", "U", "S", "L", "Encoder", "Decoder", "VariationalEncoderBlock", "VariationalDecoderBlock",
                 "Discriminator", "Autoencoder", "DoubleAutoencoder", "DoubleVariationalAutoencoder", "VariationalAutoencoder",
                 "AEGAN", "VAEGAN", "CycleAE", "CycleVAE", "CycleAEGAN", "CycleVAEGAN"):
        assert hasattr(N, name), name
    m = N.CaSb(3, 64, 7)
    assert tuple(m.conv.weight.shape) == (64, 3, 7, 7) and m.padding == 3 and m.use_norm
    with pytest.raises(NotImplementedError):
        N.CaSb(3, 8, 3, activation="Swish")
    v = N.VariationalEncoderBlock(1024, 32)
    assert tuple(v.logvarConv[1].conv.weight.shape) == (32, 32, 3, 3)
    for meth in ("configure_optimizers", "configure_loss", "training_step", "validation_step", "save_optimizer_states",
                 "load_optimizer_states"):
        assert callable(getattr(N.CycleVAEGAN, meth))
    from vcg_b200 import Losses
    for name in ("TranslationLoss", "CycleConsistencyLoss", "IdentityLoss", "GANLossGenerator", "GANLossDiscriminator",
                 "KLDivergenceLoss"):
        assert hasattr(Losses, name)


def test_unconfigured_errors_match_reference(N):
    m = N.VariationalAutoencoder()
    with pytest.raises(ValueError, match="Optimizer has not been configured"):
        m.training_step({"x": torch.zeros(1, 3, 256, 256), "y": torch.zeros(1, 3, 256, 256)})
    with pytest.raises(ValueError, match="Optimizer has not been configured"):
        m.save_optimizer_states()
    g = N.CycleVAEGAN(paired=False)
    with pytest.raises(ValueError, match="Optimizers have not been configured"):
        g.training_step({"x": torch.zeros(1, 3, 256, 256), "y": torch.zeros(1, 3, 256, 256)})


def test_no_cpu_fallback(N):
    """a CPU tensor must fail loudly, never run a PyTorch fallback."""
    enc = N.Encoder()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 3, 64, 64))
    from vcg_b200 import Losses
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Losses.TranslationLoss()(torch.zeros(4), torch.zeros(4))


def test_plan_graph_of_the_vae_generator(N):
    """the emitted DAG: 18 conv nodes + the bottleneck, activation shapes of SURVEY.md 8(a)."""
    from vcg_b200 import lib as L
    from vcg_b200.plan import ConvNode, PlanBuilder, ReparamNode
    vae = N.VariationalAutoencoder(latent_dim=64)
    b = PlanBuilder(2)
    x = b.input(3, 256, 256)
    out, node = vae.emit(b, x)
    convs = [n for n in b.nodes if isinstance(n, ConvNode)]
    assert len(convs) == 18 and sum(isinstance(n, ReparamNode) for n in b.nodes) == 1
    shapes = [(n.out_acts[0].c_log, n.out_acts[0].h, n.out_acts[0].w) for n in convs]
    assert shapes[:7] == [(64, 256, 256), (128, 128, 128), (256, 64, 64), (512, 32, 32), (1024, 16, 16),
                          (1024, 16, 16), (1024, 16, 16)]
    assert shapes[7:11] == [(64, 16, 16), (64, 16, 16), (64, 16, 16), (1024, 16, 16)]
    assert shapes[-1] == (3, 256, 256) and (out.c, out.c_log) == (8, 3)
    modes = [n.mode for n in convs]
    assert modes[1:5] == [L.MODE_UNSHUFFLE] * 4 and modes[13:17] == [L.MODE_SHUFFLE] * 4
    # ReLU BEFORE InstanceNorm in D/U/R.conv1 (Networks.py:94,111,129): fused in the conv epilogue
    assert convs[1].pre_act == L.ACT_RELU and convs[1].out_acts[0].norm and convs[1].out_acts[0].act == L.ACT_NONE
    # first layer: conv -> IN -> ReLU (Networks.py:77-80): activation applied after the norm
    assert convs[0].pre_act == L.ACT_NONE and convs[0].out_acts[0].act == L.ACT_RELU
    # the encoder's R block output feeds mu and logvar[0] convs (shared padded input) and carries a residual
    enc_out = convs[7].inp
    assert enc_out.res is not None and convs[8].inp is enc_out
    # physical channel orders: PixelUnshuffle conv reads 4*C channels, PixelShuffle conv C/4
    assert convs[1].spec.cin_phys == 256 and convs[13].spec.cin_phys == 256 and convs[16].spec.cin_phys == 32


def test_discriminator_plan_and_size_check(N):
    from vcg_b200.plan import PlanBuilder
    d = N.Discriminator()
    b = PlanBuilder(1)
    d.emit(b, b.input(3, 256, 256))
    assert [n.spec.pkh for n in b.nodes[:4]] == [2, 2, 2, 2]          # k4 s2 convs run as k2 s1 over space-to-depth
    assert b.nodes[0].spec.cin_phys == 32 and b.nodes[1].spec.cin_phys == 256
    with pytest.raises(ValueError, match="256x256"):
        d.emit(PlanBuilder(1), PlanBuilder(1).input(3, 128, 128))


def test_cli_aliases_and_defaults(vcg):
    from vcg_b200 import train
    for alias, full in (("ae", "autoencoder"), ("vae", "vae"), ("aegan", "aegan"), ("vae_gan", "vaegan"),
                        ("cycle_vae", "cyclevae"), ("vae_cyclegan", "cyclevaegan")):
        assert train.canonical_architecture(alias) == full
    with pytest.raises(ValueError):
        train.canonical_architecture("resnet")
    a = train.build_parser().parse_args([])
    assert (a.batch_size, a.lr, a.lambda_kl, a.lambda_gan, a.lambda_identity, a.lambda_cycle, a.lambda_recon, a.paired,
            a.latent_dim) == (5, 2e-4, 1e-5, 1.0, 5.0, 10.0, 1.0, False, 64)
    m = train.create_model("vae_cyclegan", paired=False, latent_dim=32)
    assert type(m).__name__ == "CycleVAEGAN" and m.G.latent_dim == 32 and not m.paired


def test_fused_adam_keeps_torch_adam_interface(N):
    from vcg_b200.optim import FusedAdam
    m = N.Discriminator()
    opt = FusedAdam(m.parameters(), lr=2e-4, betas=(0.5, 0.999))
    assert isinstance(opt, torch.optim.Adam)
    g = opt.param_groups[0]
    assert g["lr"] == 2e-4 and g["betas"] == (0.5, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 0
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"}
    torch.optim.Adam(m.parameters(), lr=2e-4, betas=(0.5, 0.999)).load_state_dict(sd)
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()


def test_gradient_buckets_partition_the_parameters_in_backward_order(N):
    """FusedAdam.set_owners / _build_buckets: every parameter in exactly one bucket, one owner per bucket, buckets of
    an owner in backward-completion order (reverse definition order), at most BUCKET_PARAMS parameters unless a single
    layer is larger, holders listed with every bucket (what plan.run_backward's contributed() calls count down)."""
    from vcg_b200 import optim
    m = N.CycleVAEGAN(paired=False)
    m.configure_optimizers(lr=2e-4)
    for opt, owners in ((m.optimizer_G, [m.F, m.G]), (m.optimizer_D, [m.DX, m.DY])):
        opt._build_buckets()
        seen = []
        for b in opt._buckets:
            assert b.owner is not None and b.holders, "every parameter of these models belongs to an announced owner"
            n = sum(p.numel() for p in b.params)
            assert n <= optim.BUCKET_PARAMS or len(b.holders) == 1
            assert {id(p) for h in b.holders for p in h.parameters(recurse=False)} == {id(p) for p in b.params}
            seen += [id(p) for p in b.params]
        assert sorted(seen) == sorted(id(p) for g in opt.param_groups for p in g["params"]) and len(seen) == len(set(seen))
        assert [b.owner for b in opt._buckets] == sorted((b.owner for b in opt._buckets), key=owners.index)
        for owner in owners:
            order = [id(h) for b in opt._buckets if b.owner is owner for h in b.holders]
            defined = [id(h) for h in owner.modules() if getattr(h, "_vcg_holder", False)]
            assert order == defined[::-1]
    assert len([b for b in m.optimizer_G._buckets if b.owner is m.G]) == 3 and len(m.optimizer_D._buckets) == 2
