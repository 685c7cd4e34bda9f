"""CPU oracle: a functional restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or the CPU baseline -- never as the thing measured as "ours" or shipped.

What it restates (all file:line relative to the reference checkout):
  * atomic blocks CaSb/D/R/U/S/L ........ Networks.py:57-149
  * Encoder / Decoder / VEB / VDB ........ Networks.py:154-237
  * Discriminator + spectral-norm head ... Networks.py:240-269 and
    torch/nn/utils/spectral_norm.py:92-114 (power iteration), 125-130
  * the six atomic losses ................ Losses.py:14-121
  * the eight composite training steps ... Networks.py:334-384 (Autoencoder),
    918-953 (VAE), 1068-1136 (AEGAN), 1254-1308 (VAEGAN), 1397-1439 (CycleAE),
    1525-1572 (CycleVAE), 1712-1807 (CycleAEGAN), 1973-2078 (CycleVAEGAN)
  * the two shared-encoder pretraining models ... Networks.py:415-605 (DoubleAutoencoder),
    608-852 (DoubleVariationalAutoencoder)
  * the constructor RNG order (nested ``apply(_init_weights)`` re-draws)

``emulate_bf16()`` additionally restates WHERE the sm_100a kernels round to bfloat16 (conv operands and
stored activations / gradients; accumulation, statistics, losses and Adam stay fp32), so that the
tensor-core path can be held to a tight tolerance end to end instead of to the reference's own bf16
noise: see the "bf16 emulation" section below.

The arithmetic itself lives in a third-party dependency of the reference
(``requirements.txt:1``: ``torch>=2.0.0``, unpinned; 2.11.0+cu128 in this image),
so this port calls the same ATen ops (conv2d, reflection pad, instance_norm,
pixel (un)shuffle, l1/mse, Adam) in plain functional form over a flat
``OrderedDict`` of tensors keyed exactly like the reference ``state_dict``.

Pinning: the reference ships no tests or golden vectors,
so parity is pinned by running the real reference in the build container:
``oracle/make_golden.py`` imports ``/root/reference`` and checks this port
against it (init bit-exact, metrics bit-exact on CPU) and writes the fixtures in
``tests/golden/``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# structure tables
# --------------------------------------------------------------------------- #
# (name, cin, cout, k) in module-definition order, i.e. the order nn.Module.apply
# and state_dict() visit them.
ENCODER_CONVS = [
    ("model.0.conv", 3, 64, 7),
    ("model.1.conv", 256, 128, 3),
    ("model.2.conv", 512, 256, 3),
    ("model.3.conv", 1024, 512, 3),
    ("model.4.conv", 2048, 1024, 3),
    ("model.5.conv1", 1024, 1024, 3),
    ("model.5.conv2", 1024, 1024, 3),
]
DECODER_CONVS = [
    ("model.0.conv1", 1024, 1024, 3),
    ("model.0.conv2", 1024, 1024, 3),
    ("model.1.conv", 256, 512, 3),
    ("model.2.conv", 128, 256, 3),
    ("model.3.conv", 64, 128, 3),
    ("model.4.conv", 32, 64, 3),
    ("model.5.conv", 64, 3, 7),
]
DISC_CONVS = [
    ("model.0.conv", 3, 64, 4),
    ("model.1.conv", 64, 128, 4),
    ("model.2.conv", 128, 256, 4),
    ("model.3.conv", 256, 512, 4),
]


def veb_convs(latent):
    return [
        ("muConv.conv", 1024, latent, 3),
        ("logvarConv.0.conv", 1024, latent, 3),
        ("logvarConv.1.conv", latent, latent, 3),
    ]


def vdb_convs(latent):
    return [("conv.conv", latent, 1024, 3)]


ARCHS = ("autoencoder", "vae", "aegan", "vaegan", "cycleae", "cyclevae",
         "cycleaegan", "cyclevaegan")
DOUBLE_ARCHS = ("doubleae", "doublevae")
ALIASES = {"ae": "autoencoder", "vae_gan": "vaegan", "cycle_vae": "cyclevae",
           "vae_cyclegan": "cyclevaegan"}


def canonical_arch(name):
    name = ALIASES.get(name, name)
    if name not in ARCHS and name not in DOUBLE_ARCHS:
        raise ValueError(f"unknown architecture {name!r}")
    return name


# --------------------------------------------------------------------------- #
# constructor RNG replay
# --------------------------------------------------------------------------- #
def _default_conv_init(state, key, cin, cout, k, dtype):
    """nn.Conv2d.reset_parameters (torch/nn/modules/conv.py): kaiming_uniform_(a=sqrt(5))
    on the weight, then uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) on the bias."""
    w = torch.empty(cout, cin, k, k, dtype=dtype)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(cin * k * k)
    b = torch.empty(cout, dtype=dtype)
    torch.nn.init.uniform_(b, -bound, bound)
    state[key + ".weight"] = w
    state[key + ".bias"] = b


def _kaiming(state, key, nonlinearity="relu", a=0.0):
    """The `_init_weights` hook every class applies (e.g. Networks.py:168-178)."""
    wkey = key + ".weight" if key + ".weight" in state else key + ".weight_orig"
    torch.nn.init.kaiming_normal_(state[wkey], a=a, mode="fan_out", nonlinearity=nonlinearity)
    torch.nn.init.zeros_(state[key + ".bias"])


def _build_convnet(state, prefix, convs, dtype, apply_init=True):
    for name, cin, cout, k in convs:
        _default_conv_init(state, prefix + name, cin, cout, k, dtype)
    if apply_init:
        for name, *_ in convs:
            _kaiming(state, prefix + name)


def _build_discriminator(state, prefix, dtype):
    """Networks.py:240-265.  spectral_norm registers weight_orig + buffers u, v drawn with
    normal_(0,1) then normalised (torch/nn/utils/spectral_norm.py:160-180)."""
    for name, cin, cout, k in DISC_CONVS:
        _default_conv_init(state, prefix + name, cin, cout, k, dtype)
    tmp = OrderedDict()
    _default_conv_init(tmp, "h", 512, 1, 16, dtype)
    # state_dict order of a spectral-normed conv: bias, weight_orig, weight_u, weight_v
    state[prefix + "model.4.bias"] = tmp["h.bias"]
    state[prefix + "model.4.weight_orig"] = tmp["h.weight"]
    u = F.normalize(torch.empty(1, dtype=dtype).normal_(0, 1), dim=0, eps=1e-12)
    v = F.normalize(torch.empty(512 * 16 * 16, dtype=dtype).normal_(0, 1), dim=0, eps=1e-12)
    state[prefix + "model.4.weight_u"] = u
    state[prefix + "model.4.weight_v"] = v
    for name, *_ in DISC_CONVS:
        _kaiming(state, prefix + name, "leaky_relu", 0.2)
    _kaiming(state, prefix + "model.4", "leaky_relu", 0.2)


def _disc_keys(prefix):
    return [prefix + n for n, *_ in DISC_CONVS] + [prefix + "model.4"]


def _build_ae(state, prefix, dtype):
    """Autoencoder.__init__ (Networks.py:277-288): Encoder(), Decoder(), then apply again."""
    _build_convnet(state, prefix + "encoder.", ENCODER_CONVS, dtype)
    _build_convnet(state, prefix + "decoder.", DECODER_CONVS, dtype)
    for k in _ae_keys(prefix):
        _kaiming(state, k)


def _ae_keys(prefix):
    return ([prefix + "encoder." + n for n, *_ in ENCODER_CONVS] +
            [prefix + "decoder." + n for n, *_ in DECODER_CONVS])


def _build_vae(state, prefix, latent, dtype):
    """VariationalAutoencoder.__init__ (Networks.py:856-871)."""
    _build_convnet(state, prefix + "encoder.", ENCODER_CONVS, dtype)
    _build_convnet(state, prefix + "variational_encoder_block.", veb_convs(latent), dtype, False)
    _build_convnet(state, prefix + "variational_decoder_block.", vdb_convs(latent), dtype, False)
    _build_convnet(state, prefix + "decoder.", DECODER_CONVS, dtype)
    for k in _vae_keys(prefix, latent):
        _kaiming(state, k)


def _vae_keys(prefix, latent):
    return ([prefix + "encoder." + n for n, *_ in ENCODER_CONVS] +
            [prefix + "variational_encoder_block." + n for n, *_ in veb_convs(latent)] +
            [prefix + "variational_decoder_block." + n for n, *_ in vdb_convs(latent)] +
            [prefix + "decoder." + n for n, *_ in DECODER_CONVS])


def init_state(arch, latent_dim=64, dtype=torch.float32):
    """Replays the reference constructors' RNG consumption so that
    ``torch.manual_seed(s); init_state(arch)`` equals
    ``torch.manual_seed(s); Networks.<Class>().state_dict()`` bit for bit."""
    arch = canonical_arch(arch)
    st = OrderedDict()
    if arch == "autoencoder":
        _build_ae(st, "", dtype)
    elif arch == "vae":
        _build_vae(st, "", latent_dim, dtype)
    elif arch == "aegan":                       # Networks.py:992-999
        _build_ae(st, "G.", dtype)
        _build_discriminator(st, "D.", dtype)
        for k in _ae_keys("G.") + _disc_keys("D."):
            _kaiming(st, k)
    elif arch == "vaegan":                      # Networks.py:1192-1196 (no outer apply)
        _build_vae(st, "G.", latent_dim, dtype)
        _build_discriminator(st, "D.", dtype)
    elif arch == "cycleae":                     # Networks.py:1351-1355
        _build_ae(st, "F.", dtype)
        _build_ae(st, "G.", dtype)
    elif arch == "cyclevae":                    # Networks.py:1483-1487
        _build_vae(st, "F.", latent_dim, dtype)
        _build_vae(st, "G.", latent_dim, dtype)
    elif arch == "cycleaegan":                  # Networks.py:1619-1628
        _build_ae(st, "F.", dtype)
        _build_ae(st, "G.", dtype)
        _build_discriminator(st, "DX.", dtype)
        _build_discriminator(st, "DY.", dtype)
        for k in _ae_keys("F.") + _ae_keys("G.") + _disc_keys("DX.") + _disc_keys("DY."):
            _kaiming(st, k)
    elif arch == "doubleae":                    # Networks.py:434-440 (no outer apply)
        _build_convnet(st, "encoder.", ENCODER_CONVS, dtype)
        _build_convnet(st, "decoder_A.", DECODER_CONVS, dtype)
        _build_convnet(st, "decoder_B.", DECODER_CONVS, dtype)
    elif arch == "doublevae":                   # Networks.py:626-645 (outer apply re-draws everything)
        _build_convnet(st, "encoder.", ENCODER_CONVS, dtype)
        for w in "AB":
            _build_convnet(st, f"vae_encoder_block_{w}.", veb_convs(latent_dim), dtype, False)
        for w in "AB":
            _build_convnet(st, f"vae_decoder_block_{w}.", vdb_convs(latent_dim), dtype, False)
        _build_convnet(st, "decoder_A.", DECODER_CONVS, dtype)
        _build_convnet(st, "decoder_B.", DECODER_CONVS, dtype)
        for k in ([f"encoder.{n}" for n, *_ in ENCODER_CONVS] +
                  [f"vae_encoder_block_{w}.{n}" for w in "AB" for n, *_ in veb_convs(latent_dim)] +
                  [f"vae_decoder_block_{w}.{n}" for w in "AB" for n, *_ in vdb_convs(latent_dim)] +
                  [f"decoder_{w}.{n}" for w in "AB" for n, *_ in DECODER_CONVS]):
            _kaiming(st, k)
    elif arch == "cyclevaegan":                 # Networks.py:1874-1883
        _build_vae(st, "F.", latent_dim, dtype)
        _build_vae(st, "G.", latent_dim, dtype)
        _build_discriminator(st, "DX.", dtype)
        _build_discriminator(st, "DY.", dtype)
        for k in (_vae_keys("F.", latent_dim) + _vae_keys("G.", latent_dim) +
                  _disc_keys("DX.") + _disc_keys("DY.")):
            _kaiming(st, k)
    return st


def is_buffer(key):
    return key.endswith("weight_u") or key.endswith("weight_v")


def param_keys(state, prefixes):
    """Parameter keys (buffers excluded) in nn.Module.parameters() order.  For a
    spectral-normed conv that order is bias, weight_orig -- the state_dict order."""
    return [k for k in state if any(k.startswith(p) for p in prefixes) and not is_buffer(k)]


# --------------------------------------------------------------------------- #
# bf16 emulation: where the sm_100a tensor-core path rounds
# --------------------------------------------------------------------------- #
# In bf16 mode the kernels (vae-cyclegan-implementation_b200/csrc) keep fp32 master weights, fp32
# accumulators, fp32 InstanceNorm statistics, fp32 losses and fp32 Adam, and round to bfloat16 exactly here:
#   (r1) a network's NCHW fp32 input when it is packed to NHWC (pack_nchw);
#   (r2) each filter when it is packed for the GEMM (wpack_multi): forward value only, the weight gradient
#        accumulates and stays fp32;
#   (r3) each conv output after bias and the fused pre-norm activation, when the epilogue stores it -- except the
#        final image layer and the mu / second logvar conv, which are stored in fp32;
#   (r4) each conv INPUT after normalisation / activation / residual, when the transform pass stores it (xform_fwd);
#        z = mu + eps * std likewise;
#   and the gradients stored at the same places: dY of every conv (r3, also of the fp32-stored outputs), the
#   padded-input gradient dXp written by the data-gradient GEMM (after the pad, r5), the gathered gradient g between
#   the activation derivative and the InstanceNorm backward (r6), and the packed gradient of every network output.
# `emulate_bf16()` switches the blocks below to insert those roundings (straight-through: the forward value and /
# or the passing gradient is rounded to the nearest bfloat16, ties to even, like cvt.rn.bf16.f32).
_EMU = [False]


class emulate_bf16:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.prev, _EMU[0] = _EMU[0], self.on

    def __exit__(self, *exc):
        _EMU[0] = self.prev


def _bf(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return _bf(x) if fwd else x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return (_bf(g) if ctx.bwd else g), None, None


def r_both(x):       # stored activation: value and its gradient are bf16 tensors
    return _Round.apply(x, True, True) if _EMU[0] else x


def r_fwd(x):        # packed filter
    return _Round.apply(x, True, False) if _EMU[0] else x


def r_bwd(x):        # fp32-stored value whose gradient is a bf16 tensor
    return _Round.apply(x, False, True) if _EMU[0] else x


# --------------------------------------------------------------------------- #
# forward blocks
# --------------------------------------------------------------------------- #
def conv_reflect(x, w, b, stride=1, pad=1):
    """nn.Conv2d(padding_mode='reflect') = F.conv2d(F.pad(x, reflect), w, b, stride, 0)
    (torch/nn/modules/conv.py:534-550)."""
    if pad:
        x = r_bwd(F.pad(x, (pad, pad, pad, pad), mode="reflect"))                  # (r5)
    return F.conv2d(x, r_fwd(w), b, stride)                                         # (r2)


def inorm(x):
    """nn.InstanceNorm2d defaults: affine=False, no running stats, eps=1e-5, biased var."""
    return r_bwd(F.instance_norm(x, eps=1e-5))                                      # (r6)


def _cw(P, key):
    return P[key + ".weight"], P[key + ".bias"]


def encoder(P, pre, x):
    """Networks.py:154-181 with blocks 57-116."""
    x = r_both(x)                                                                   # (r1)
    h = r_both(conv_reflect(x, *_cw(P, pre + "model.0.conv"), 1, 3))                # CaSb: conv, IN, ReLU   (r3)
    h = r_both(F.relu(inorm(h)))                                                    # (r4)
    for i in (1, 2, 3, 4):                                                          # D: unshuffle, conv, ReLU, IN
        h = r_both(F.relu(conv_reflect(F.pixel_unshuffle(h, 2), *_cw(P, pre + f"model.{i}.conv"))))
        h = r_both(inorm(h))
    return res_block(P, pre + "model.5.", h)


def res_block(P, pre, x):
    """R (Networks.py:98-116): conv1, ReLU, IN, conv2, IN, + residual."""
    h = r_both(inorm(r_both(F.relu(conv_reflect(x, *_cw(P, pre + "conv1"))))))
    h = inorm(r_both(conv_reflect(h, *_cw(P, pre + "conv2"))))
    return r_both(h + x)


def decoder(P, pre, z):
    """Networks.py:183-199."""
    h = res_block(P, pre + "model.0.", z)
    for i in (1, 2, 3, 4):                                                          # U: shuffle, conv, ReLU, IN
        h = r_both(F.relu(conv_reflect(F.pixel_shuffle(h, 2), *_cw(P, pre + f"model.{i}.conv"))))
        h = r_both(inorm(h))
    return r_bwd(conv_reflect(h, *_cw(P, pre + "model.5.conv"), 1, 3))              # CaSb Identity, no norm; fp32 store


def ae_forward(P, pre, x, dec="decoder."):
    return decoder(P, pre + dec, encoder(P, pre + "encoder.", x))


def veb(P, pre, h, eps=None):
    """VariationalEncoderBlock.forward (Networks.py:219-227)."""
    mu = r_bwd(conv_reflect(h, *_cw(P, pre + "muConv.conv")))                       # fp32 store
    lv = r_both(conv_reflect(h, *_cw(P, pre + "logvarConv.0.conv")))
    lv = r_bwd(conv_reflect(lv, *_cw(P, pre + "logvarConv.1.conv")))                # fp32 store
    lv = torch.clamp(lv, min=-10, max=10)
    std = torch.exp(0.5 * lv)
    if eps is None:
        eps = torch.randn_like(std)
    elif callable(eps):
        eps = eps(std)
    return r_both(mu + eps * std), mu, lv


def vae_forward(P, pre, x, eps=None, veb_pre="variational_encoder_block.", vdb_pre="variational_decoder_block.",
                dec="decoder."):
    """VariationalAutoencoder.forward (Networks.py:885-890) -> (Gx, mu, logvar)."""
    h = encoder(P, pre + "encoder.", x)
    z, mu, lv = veb(P, pre + veb_pre, h, eps)
    h = r_both(conv_reflect(z, *_cw(P, pre + vdb_pre + "conv.conv")))
    return decoder(P, pre + dec, h), mu, lv


def spectral_weight(P, pre, training=True):
    """torch/nn/utils/spectral_norm.py:92-114: one power iteration per training forward,
    u and v updated in place (no grad), sigma = u . (W v), weight = W / sigma."""
    w = P[pre + "weight_orig"]
    u, v = P[pre + "weight_u"], P[pre + "weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def discriminator(P, pre, x, training=True):
    """Networks.py:240-269."""
    x = r_both(x)
    h = r_both(F.leaky_relu(conv_reflect(x, *_cw(P, pre + "model.0.conv"), 2, 1), 0.2))
    for i in (1, 2, 3):
        h = r_both(conv_reflect(h, *_cw(P, pre + f"model.{i}.conv"), 2, 1))
        h = r_both(F.leaky_relu(inorm(h), 0.2))
    w = spectral_weight(P, pre + "model.4.", training)                              # fp32 head: a dot product, not a GEMM
    return F.conv2d(h, w, P[pre + "model.4.bias"]).view(-1, 1).squeeze(1)


# --------------------------------------------------------------------------- #
# losses (Losses.py:14-121)
# --------------------------------------------------------------------------- #
def l1(a, b):
    return F.l1_loss(a, b)


def cycle_loss(x, y, FGx, GFy):
    return l1(FGx, x) + l1(GFy, y)


def identity_loss(x, y, Fx, Gy):
    return l1(Fx, x) + l1(Gy, y)


def gan_loss_gen(d_real, d_fake):
    real = F.mse_loss(d_real, torch.zeros_like(d_real))
    fake = F.mse_loss(d_fake, torch.ones_like(d_fake))
    return real + fake, real, fake


def gan_loss_disc(d_real, d_fake):
    real = F.mse_loss(d_real, torch.full_like(d_real, 1))
    fake = F.mse_loss(d_fake, torch.full_like(d_fake, 0))
    return real + fake, real, fake


def kl_loss(mu, logvar):
    lv = torch.clamp(logvar, min=-10, max=10)
    return -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())


# --------------------------------------------------------------------------- #
# composite models with training_step
# --------------------------------------------------------------------------- #
DEFAULT_LAMBDAS = dict(lambda_kl=1e-5, lambda_gan=1.0, lambda_identity=5.0,
                       lambda_cycle=10.0, lambda_recon=1.0)


class RefModel:
    """Flat-state restatement of one reference composite class.

    ``state`` has the reference's ``state_dict`` keys.  ``eps_source`` (optional) is a
    callable ``eps_source(like)`` replacing ``torch.randn_like`` so that several precisions
    / implementations can be fed identical noise; it is called in the reference's order.
    """

    def __init__(self, arch, latent_dim=64, paired=False, state=None, dtype=torch.float32,
                 lr=2e-4, betas=(0.5, 0.999), lambdas=None, eps_source=None, emulate_bf16=False, device=None):
        """emulate_bf16: round like the sm_100a tensor-core path (see "bf16 emulation" above); device: where the
        tensors live (the GPU parity tests run this checker on the box's GPU in fp32 at the BASELINE batch sizes)."""
        self.arch = canonical_arch(arch)
        self.emulate = bool(emulate_bf16)
        self.latent_dim = latent_dim
        self.paired = paired
        self.lam = dict(DEFAULT_LAMBDAS)
        self.lam.update(lambdas or {})
        st = state if state is not None else init_state(self.arch, latent_dim, dtype)
        self.P = OrderedDict()
        for k, v in st.items():
            t = v.detach().to(dtype).clone()
            if device is not None:
                t = t.to(device)
            if not is_buffer(k):
                t.requires_grad_(True)
            self.P[k] = t
        self.eps_source = eps_source
        self.training = True
        a = self.arch
        if a in ("autoencoder", "vae", "cycleae", "cyclevae", "doubleae", "doublevae"):
            # Adam(self.parameters()) -- Networks.py:312, 894, 1372, 1498
            self.opt_G = torch.optim.Adam([self.P[k] for k in param_keys(self.P, [""])], lr=lr, betas=betas)
            self.opt_D = None
        elif a in ("aegan", "vaegan"):           # Networks.py:1032-1033, 1212-1213
            self.opt_G = torch.optim.Adam([self.P[k] for k in param_keys(self.P, ["G."])], lr=lr, betas=betas)
            self.opt_D = torch.optim.Adam([self.P[k] for k in param_keys(self.P, ["D."])], lr=lr, betas=betas)
        else:                                    # Networks.py:1669-1676, 1928-1935: F then G; DX then DY
            self.opt_G = torch.optim.Adam([self.P[k] for k in param_keys(self.P, ["F."])] +
                                          [self.P[k] for k in param_keys(self.P, ["G."])], lr=lr, betas=betas)
            self.opt_D = torch.optim.Adam([self.P[k] for k in param_keys(self.P, ["DX."])] +
                                          [self.P[k] for k in param_keys(self.P, ["DY."])], lr=lr, betas=betas)

    # -- helpers ------------------------------------------------------------
    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.P.items())

    def _vae(self, pre, x):
        # eps_source(std) is called exactly where the reference calls randn_like(std)
        return vae_forward(self.P, pre, x, self.eps_source)

    def _gen(self, pre, x):
        if self.arch in ("vae", "vaegan", "cyclevae", "cyclevaegan"):
            return self._vae(pre, x)
        return (ae_forward(self.P, pre, x),)

    def _disc(self, pre, x):
        return discriminator(self.P, pre, x, self.training)

    # -- forward with the reference's return orders ---------------------------
    def forward(self, x, y=None):
        with emulate_bf16(self.emulate):
            return self._forward(x, y)

    def _forward(self, x, y=None):
        a = self.arch
        if a == "doubleae":                                # Networks.py:447-466
            return ae_forward(self.P, "", x, "decoder_A."), ae_forward(self.P, "", y, "decoder_B.")
        if a == "doublevae":                               # Networks.py:661-685 (noise drawn for A, then for B)
            Gx, mu_x, lv_x = vae_forward(self.P, "", x, self.eps_source, "vae_encoder_block_A.",
                                         "vae_decoder_block_A.", "decoder_A.")
            Gy, mu_y, lv_y = vae_forward(self.P, "", y, self.eps_source, "vae_encoder_block_B.",
                                         "vae_decoder_block_B.", "decoder_B.")
            return Gx, Gy, mu_x, lv_x, mu_y, lv_y
        if a == "autoencoder":
            return ae_forward(self.P, "", x)
        if a == "vae":
            return self._vae("", x)
        if a == "aegan":                                   # Networks.py:1023-1028
            Gx = ae_forward(self.P, "G.", x)
            Gy = ae_forward(self.P, "G.", y)
            return Gx, Gy, self._disc("D.", Gx), self._disc("D.", y)
        if a == "vaegan":                                  # Networks.py:1203-1208
            Gx, mu, lv = self._vae("G.", x)
            Gy, mu_y, lv_y = self._vae("G.", y)
            return Gx, mu, lv, Gy, mu_y, lv_y, self._disc("D.", Gx), self._disc("D.", y)
        if a == "cycleae":                                 # Networks.py:1363-1368
            Gx = ae_forward(self.P, "G.", x)
            FGx = ae_forward(self.P, "F.", Gx)
            Fy = ae_forward(self.P, "F.", y)
            return Gx, FGx, Fy, ae_forward(self.P, "G.", Fy)
        if a == "cyclevae":                                # Networks.py:1489-1494
            Gx, mu_x, lv_x = self._vae("G.", x)
            FGx, mu_FGx, lv_FGx = self._vae("F.", Gx)
            Fy, mu_y, lv_y = self._vae("F.", y)
            GFy, mu_GFy, lv_GFy = self._vae("G.", Fy)
            return (Gx, FGx, Fy, GFy, mu_x, lv_x, mu_FGx, lv_FGx, mu_y, lv_y, mu_GFy, lv_GFy)
        if a == "cycleaegan":                              # Networks.py:1654-1665
            Gx = ae_forward(self.P, "G.", x)
            Gy = ae_forward(self.P, "G.", y)
            FGx = ae_forward(self.P, "F.", Gx)
            Fy = ae_forward(self.P, "F.", y)
            Fx = ae_forward(self.P, "F.", x)
            GFy = ae_forward(self.P, "G.", Fy)
            return (Gx, FGx, Fy, GFy, self._disc("DY.", Gx), self._disc("DX.", Fy),
                    self._disc("DX.", x), self._disc("DY.", y), Gy, Fx)
        # cyclevaegan: Networks.py:1909-1924
        Gx, mu_x, lv_x = self._vae("G.", x)
        Gy, _, _ = self._vae("G.", y)
        FGx, mu_FGx, lv_FGx = self._vae("F.", Gx)
        Fy, mu_y, lv_y = self._vae("F.", y)
        Fx, _, _ = self._vae("F.", x)
        GFy, mu_GFy, lv_GFy = self._vae("G.", Fy)
        return (Gx, FGx, Fy, GFy, mu_x, lv_x, mu_FGx, lv_FGx, mu_y, lv_y, mu_GFy, lv_GFy,
                self._disc("DY.", Gx), self._disc("DX.", Fy), self._disc("DX.", x),
                self._disc("DY.", y), Gy, Fx)

    # -- training steps -------------------------------------------------------
    def training_step(self, batch):
        with emulate_bf16(self.emulate):
            return getattr(self, "_step_" + self.arch)(batch["x"], batch["y"])

    def translate(self, x, to):
        """Double models: translate_A_to_B (to='B') / translate_B_to_A (to='A'), Networks.py:468-476, 687-699."""
        with emulate_bf16(self.emulate):
            if self.arch == "doubleae":
                return ae_forward(self.P, "", x, f"decoder_{to}.")
            return vae_forward(self.P, "", x, self.eps_source, f"vae_encoder_block_{to}.", f"vae_decoder_block_{to}.",
                               f"decoder_{to}.")[0]

    def _step_doubleae(self, x, y):                        # Networks.py:505-546
        Gx, Gy = self._forward(x, y)
        la, lb = l1(Gx, x), l1(Gy, y)
        total = la + lb
        self.opt_G.zero_grad()
        total.backward()
        self.opt_G.step()
        return {"G_loss": total.item(), "loss_recon_A": la.item(), "loss_recon_B": lb.item(), "total_loss": total.item()}

    def _step_doublevae(self, x, y):                       # Networks.py:757-799
        Gx, Gy, mu_x, lv_x, mu_y, lv_y = self._forward(x, y)
        la, lb = l1(Gx, x), l1(Gy, y)
        ka, kb = kl_loss(mu_x, lv_x), kl_loss(mu_y, lv_y)
        lk = ka + kb
        total = la + lb + self.lam["lambda_kl"] * lk
        self.opt_G.zero_grad()
        total.backward()
        self.opt_G.step()
        return {"G_loss": total.item(), "loss_recon_A": la.item(), "loss_recon_B": lb.item(), "loss_kl": lk.item(),
                "loss_kl_A": ka.item(), "loss_kl_B": kb.item(), "total_loss": total.item()}

    def _step_autoencoder(self, x, y):                     # Networks.py:334-384
        loss = l1(self.forward(x), y)
        self.opt_G.zero_grad()
        loss.backward()
        self.opt_G.step()
        v = loss.item()
        return {"G_loss": v, "loss_trans": v, "total_loss": v}

    def _step_vae(self, x, y):                             # Networks.py:918-953
        out, mu, lv = self.forward(x)
        lt, lk = l1(out, y), kl_loss(mu, lv)
        G_loss = lt + self.lam["lambda_kl"] * lk
        self.opt_G.zero_grad()
        G_loss.backward()
        self.opt_G.step()
        return {"G_loss": G_loss.item(), "loss_trans": lt.item(), "loss_kl": lk.item()}

    def _step_aegan(self, x, y):                           # Networks.py:1068-1136
        self.opt_G.zero_grad()
        Gx, Gy, DGx, Dy = self.forward(x, y)
        lt = l1(Gx, y)
        lg, lg_real, lg_fake = gan_loss_gen(Dy, DGx)
        lid = l1(Gy, y)
        G_loss = lt + self.lam["lambda_gan"] * lg + self.lam["lambda_identity"] * lid
        G_loss.backward()
        self.opt_G.step()
        self.opt_D.zero_grad()
        DGx_d = self._disc("D.", Gx.detach())
        Dy_d = self._disc("D.", y)
        D_loss, D_real, D_fake = gan_loss_disc(Dy_d, DGx_d)
        D_loss.backward()
        self.opt_D.step()
        return {"G_loss": G_loss.item(), "D_loss": D_loss.item(), "D_loss_real": D_real.item(),
                "D_loss_fake": D_fake.item(), "loss_trans": lt.item(), "loss_gan_g": lg.item(),
                "loss_identity": lid.item(), "d_y_mean": Dy_d.mean().item(),
                "d_gx_mean": DGx_d.mean().item()}

    def _step_vaegan(self, x, y):                          # Networks.py:1254-1308
        Gx, mu, lv, Gy, mu_y, lv_y, DGx, Dy = self.forward(x, y)
        lt = l1(Gx, y)
        lg, lg_real, lg_fake = gan_loss_gen(Dy, DGx)
        lid = l1(Gy, y)
        lk = kl_loss(mu, lv)
        lam = self.lam
        G_loss = (lam["lambda_recon"] * lt + lam["lambda_gan"] * lg +
                  lam["lambda_identity"] * lid + lam["lambda_kl"] * lk)
        D_loss, D_real, D_fake = gan_loss_disc(Dy, DGx.detach())   # only Dy trains D (quirk kept)
        self.opt_G.zero_grad()
        G_loss.backward(retain_graph=True)
        self.opt_G.step()
        self.opt_D.zero_grad()
        D_loss.backward()
        self.opt_D.step()
        return {"G_loss": G_loss.item(), "D_loss": D_loss.item(),
                "loss_gan_disc_real": D_real.item(), "loss_gan_disc_fake": D_fake.item(),
                "loss_trans": lt.item(), "loss_gan_real": lg_real.item(),
                "loss_gan_fake": lg_fake.item(), "loss_identity": lid.item(), "loss_kl": lk.item()}

    def _step_cycleae(self, x, y):                         # Networks.py:1397-1439
        Gx, FGx, Fy, GFy = self.forward(x, y)
        lc = cycle_loss(x, y, FGx, GFy)
        total = self.lam["lambda_cycle"] * lc
        m = {"total_loss": total.item(), "loss_cycle": lc.item(), "G_loss": total.item()}
        if self.paired:
            lt = l1(Gx, y) + l1(Fy, x)
            total = total + lt
            m.update(loss_trans=lt.item(), total_loss=total.item(), G_loss=total.item())
        self.opt_G.zero_grad()
        total.backward()
        self.opt_G.step()
        return m

    def _step_cyclevae(self, x, y):                        # Networks.py:1525-1572
        (Gx, FGx, Fy, GFy, mu_x, lv_x, mu_FGx, lv_FGx, mu_y, lv_y, mu_GFy, lv_GFy) = self.forward(x, y)
        lc = cycle_loss(x, y, FGx, GFy)
        lk = (kl_loss(mu_x, lv_x) + kl_loss(mu_FGx, lv_FGx) + kl_loss(mu_y, lv_y) + kl_loss(mu_GFy, lv_GFy))
        total = self.lam["lambda_cycle"] * lc + self.lam["lambda_kl"] * lk
        m = {"total_loss": total.item(), "loss_cycle": lc.item(), "loss_kl": lk.item(), "G_loss": total.item()}
        if self.paired:
            lt = l1(Gx, y) + l1(Fy, x)
            total = total + lt
            m.update(loss_trans=lt.item(), total_loss=total.item(), G_loss=total.item())
        self.opt_G.zero_grad()
        total.backward()
        self.opt_G.step()
        return m

    def _d_step_cycle(self, x, y, Gx, Fy):
        """Shared discriminator half of the two Cycle*GAN steps (Networks.py:1762-1785,
        2025-2051): D forwards re-run on detached fakes, LSGAN real->1 fake->0."""
        self.opt_D.zero_grad()
        DYGx = self._disc("DY.", Gx.detach())
        DXFy = self._disc("DX.", Fy.detach())
        DXx = self._disc("DX.", x)
        DYy = self._disc("DY.", y)
        ldx, dxr, dxf = gan_loss_disc(DXx, DXFy)
        ldy, dyr, dyf = gan_loss_disc(DYy, DYGx)
        D_loss = ldx + ldy
        D_loss.backward()
        self.opt_D.step()
        return D_loss, dict(D_loss=D_loss.item(), D_loss_x_real=dxr.item(), D_loss_x_fake=dxf.item(),
                            D_loss_y_real=dyr.item(), D_loss_y_fake=dyf.item(),
                            d_x_real_mean=DXx.mean().item(), d_x_fake_mean=DXFy.mean().item(),
                            d_y_real_mean=DYy.mean().item(), d_y_fake_mean=DYGx.mean().item())

    def _step_cycleaegan(self, x, y):                      # Networks.py:1712-1807
        self.opt_G.zero_grad()
        Gx, FGx, Fy, GFy, DYGx, DXFy, DXx, DYy, Gy, Fx = self.forward(x, y)
        lc = cycle_loss(x, y, FGx, GFy)
        lgx, lgx_r, lgx_f = gan_loss_gen(DXx, DXFy)
        lgy, lgy_r, lgy_f = gan_loss_gen(DYy, DYGx)
        lg = lgx + lgy
        G_loss = self.lam["lambda_cycle"] * lc + self.lam["lambda_gan"] * lg
        if self.paired:
            lid = identity_loss(x, y, Fx, Gy)
            G_loss = G_loss + self.lam["lambda_identity"] * lid
        G_loss.backward()
        self.opt_G.step()
        D_loss, dm = self._d_step_cycle(x, y, Gx, Fy)
        m = {"total_loss": G_loss.item() + D_loss.item(), "G_loss": G_loss.item()}
        m.update(dm)
        m.update(loss_cycle=lc.item(), loss_gan_g=lg.item(), loss_gan_g_x_real=lgx_r.item(),
                 loss_gan_g_x_fake=lgx_f.item(), loss_gan_g_y_real=lgy_r.item(),
                 loss_gan_g_y_fake=lgy_f.item())
        if self.paired:
            m["loss_identity"] = lid.item()
        return m

    def _step_cyclevaegan(self, x, y):                     # Networks.py:1973-2078
        self.opt_G.zero_grad()
        (Gx, FGx, Fy, GFy, mu_x, lv_x, mu_FGx, lv_FGx, mu_y, lv_y, mu_GFy, lv_GFy,
         DYGx, DXFy, DXx, DYy, Gy, Fx) = self.forward(x, y)
        lc = cycle_loss(x, y, FGx, GFy)
        lgx, lgx_r, lgx_f = gan_loss_gen(DXx, DXFy)
        lgy, lgy_r, lgy_f = gan_loss_gen(DYy, DYGx)
        lg_fake = lgx_f + lgy_f
        lk = (kl_loss(mu_x, lv_x) + kl_loss(mu_FGx, lv_FGx) + kl_loss(mu_y, lv_y) + kl_loss(mu_GFy, lv_GFy))
        lam = self.lam
        G_loss = lam["lambda_cycle"] * lc + lam["lambda_gan"] * lg_fake + lam["lambda_kl"] * lk
        if self.paired:
            lid = identity_loss(x, y, Fx, Gy)
            G_loss = G_loss + lam["lambda_identity"] * lid
        G_loss.backward()
        self.opt_G.step()
        D_loss, dm = self._d_step_cycle(x, y, Gx, Fy)
        m = {"total_loss": G_loss.item() + D_loss.item(), "G_loss": G_loss.item()}
        m.update(dm)
        m.update(loss_cycle=lc.item(), loss_gan_g=lg_fake.item(), loss_gan_g_x_real=lgx_r.item(),
                 loss_gan_g_x_fake=lgx_f.item(), loss_gan_g_y_real=lgy_r.item(),
                 loss_gan_g_y_fake=lgy_f.item(), loss_kl=lk.item())
        if self.paired:
            m["loss_identity"] = lid.item()
        return m


def synthetic_batch(batch, seed=7, same_xy=False, size=256, dtype=torch.float32):
    """x, y ~ U[0,1) from torch.Generator().manual_seed(7)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, size, size, generator=g)
    y = x if same_xy else torch.rand(batch, 3, size, size, generator=g)
    return {"x": x.to(dtype), "y": y.to(dtype)}
